#!/bin/bash
# BASELINE configs[0] pipeline (examples/cfg1_pipeline) in its four buffer modes, 1080p and 4K.
python - <<'PY'
import sys; sys.path.insert(0, ".")
from gst_plugins_rs_b200 import frames
open("/tmp/lut33.cube", "w").write(frames.cube_text_3d(33))
PY
for geom in "1920 1080" "3840 2160"; do
  for mode in pageable pool register queued; do
    examples/cfg1_pipeline /tmp/lut33.cube 300 $geom 8 0 $mode | tail -1
  done
done
