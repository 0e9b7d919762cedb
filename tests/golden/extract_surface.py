"""Extracts the element surface (names, klass, caps formats, properties with type / default /
range / mutability) of colorlut, hsvfilter and hsvdetector from the reference's
docs/plugins/gst_plugins_cache.json into tests/golden/element_surface.json.

Run in the build container, where /root/reference exists:  python tests/golden/extract_surface.py
The GPU box has no /root/reference; tests read only the committed JSON.
"""
import json
import os
import re

SRC = "/root/reference/docs/plugins/gst_plugins_cache.json"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "element_surface.json")


def formats(caps):
    m = re.search(r"format:\s*\{([^}]*)\}", caps)
    return [f.strip() for f in m.group(1).split(",")]


def main():
    cache = json.load(open(SRC))
    out = {}
    for plugin in ("colorlut", "hsv"):
        p = cache[plugin]
        for name, e in p["elements"].items():
            if name not in ("colorlut", "hsvfilter", "hsvdetector"):
                continue
            props = {}
            for pn, pd in e["properties"].items():
                if pn in ("name", "parent", "qos"):
                    continue
                props[pn] = {k: pd[k] for k in ("type", "default", "mutable", "readable", "writable")
                             if k in pd}
                for k in ("min", "max"):
                    if k in pd:
                        props[pn][k] = pd[k]
            out[name] = {
                "plugin": plugin, "filename": p["filename"], "license": p["license"],
                "gtype": e["hierarchy"][0], "parent": e["hierarchy"][1], "klass": e["klass"],
                "rank": e["rank"],
                "sink_formats": formats(e["pad-templates"]["sink"]["caps"]),
                "src_formats": formats(e["pad-templates"]["src"]["caps"]),
                "properties": props,
            }
    json.dump(out, open(DST, "w"), indent=1, sort_keys=True)
    print("wrote", DST)


if __name__ == "__main__":
    main()
