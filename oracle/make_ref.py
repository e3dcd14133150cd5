"""Oracle (test infrastructure): recipe that makes the UNMODIFIED reference model travel to the GPU box.

    python oracle/make_ref.py        # build container only: needs /root/reference

Copies `/root/reference/vit_model.py` byte for byte into `oracle/_ref/` -- a git-ignored build output (like
libvtc.so: kept out of history, shipped to the GPU box with the gpurun snapshot) -- and records its sha256 in
`oracle/_ref/MANIFEST.json`.  Nothing of it is committed.  `oracle/ref_shim.py` imports it from there when
`/root/reference` does not exist, so that `bench.py --impl reference`, the `cpu_baseline` leg and the on-box eager GPU
baseline time the reference's own code (`kind: "reference"`) instead of the oracle port (`kind: "port"`).
`__graft_entry__.build()` runs this when the reference is present."""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = "/root/reference"
DST_DIR = os.path.join(HERE, "_ref")
FILES = ("vit_model.py",)


def make() -> bool:
    if not os.path.isfile(os.path.join(SRC_DIR, FILES[0])):
        return False
    os.makedirs(DST_DIR, exist_ok=True)
    manifest = {}
    for name in FILES:
        src, dst = os.path.join(SRC_DIR, name), os.path.join(DST_DIR, name)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC_DIR, "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref written" if make() else f"{SRC_DIR} not found: nothing to do")
