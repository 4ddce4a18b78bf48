"""Copy the reference's recorded PNGs (output/page_1.png, output/pages/page_001..022.png) next to the tests so that the GPU box —
which has no /root/reference — can run the full-fixture parity test.  The copies go to tests/golden/_recorded/ (git-ignored: they
are the reference's data artefacts, only their digests are committed in fixtures.json; the directory still travels with gpurun).

Run in the build container:  python tests/golden/fetch_recorded.py
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/output"
DST = os.path.join(HERE, "_recorded")


def main():
    fx = json.load(open(os.path.join(HERE, "fixtures.json")))["fixtures"]
    os.makedirs(os.path.join(DST, "pages"), exist_ok=True)
    for name, meta in fx.items():
        src = os.path.join(REF, name)
        raw = open(src, "rb").read()
        assert hashlib.sha256(raw).hexdigest()[:16] == meta["sha_file"], name
        shutil.copyfile(src, os.path.join(DST, name))
    print(f"copied {len(fx)} recorded PNGs to {DST}")


if __name__ == "__main__":
    main()
