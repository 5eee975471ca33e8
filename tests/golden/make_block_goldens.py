#!/usr/bin/env python3
"""Writes tests/golden/block_samples.json from the reference's hard-coded sample meshes.

`Samples::block_2d_four_qua4 / qua8 / qua9?`, `block_3d_eight_hex8 / hex20` (src/mesh/samples.rs) are the fixtures the
reference uses to pin `Block::subdivide`'s cell ordering and first-touch point numbering
(src/mesh/block.rs tests :1753-2196 compare format!("{:?}", mesh) against them).  This script parses the Rust literals
(points and cells only) -- it needs /root/reference and is run in the authoring container only; the JSON is committed.
"""
import json
import re
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
NAMES = ["block_2d_four_qua4", "block_2d_four_qua8", "block_2d_four_qua9", "block_2d_four_qua12", "block_2d_four_qua16",
         "block_2d_four_qua17", "block_3d_eight_hex8", "block_3d_eight_hex20", "block_3d_eight_hex32"]


def main():
    ref = Path(sys.argv[1]) if len(sys.argv) > 1 else Path("/root/reference")
    text = (ref / "src/mesh/samples.rs").read_text()
    out = {}
    for name in NAMES:
        m = re.search(r"pub fn " + name + r"\(\) -> Mesh \{(.*?)\n    \}\n", text, re.S)
        if not m:
            continue
        body = m.group(1)
        ndim = int(re.search(r"ndim:\s*(\d+)", body).group(1))
        points = []
        for pm in re.finditer(r"Point \{ id:\s*(\d+), marker:\s*(-?\d+), coords: vec!\[([^\]]*)\]", body):
            points.append((int(pm.group(1)), [float(x) for x in pm.group(3).split(",") if x.strip()]))
        cells = []
        for cm in re.finditer(r"Cell \{ id:\s*(\d+), marker:\s*(-?\d+), kind: GeoKind::(\w+), points: vec!\[([^\]]*)\]", body):
            cells.append(dict(id=int(cm.group(1)), marker=int(cm.group(2)), kind=cm.group(3).lower(),
                              points=[int(x) for x in cm.group(4).split(",") if x.strip()]))
        assert [p[0] for p in points] == list(range(len(points)))
        out[name] = dict(ndim=ndim, coords=[p[1] for p in points], cells=cells,
                         ref="src/mesh/samples.rs fn " + name)
    (HERE / "block_samples.json").write_text(json.dumps(out))
    print({k: (len(v["coords"]), len(v["cells"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
