# Convenience targets (the driver uses __graft_entry__.build() / pytest / bench.py directly).
PY ?= python
.PHONY: build test test-gpu bench clean
build:
	$(PY) -c "import __graft_entry__ as g; g.build()"
test: build
	$(PY) -m pytest tests -x -q -m "not gpu"
test-gpu: build
	$(PY) -m pytest tests -x -q -m gpu
bench: build
	$(PY) bench.py
clean:
	$(MAKE) -C gst-plugins-rs_b200/csrc clean
	$(MAKE) -C gst-plugins-rs_b200/elements clean
	$(MAKE) -C oracle clean
	$(MAKE) -C examples clean
