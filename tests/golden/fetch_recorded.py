"""Copy the reference's recorded PNGs (output/page_1.png, output/pages/page_001..022.png) next to the tests so that the GPU box —
which has no /root/reference — can run the full-fixture parity test.  The copies go to tests/golden/_recorded/ (git-ignored: they
are the reference's data artefacts, only their digests are committed in fixtures.json; the directory still travels with gpurun).

The same run copies the recorded Gemini responses (output/pages/page_0NN.json) and the reference's call-site module
(backend/app/pipeline/pdf_extract.py with the two modules it imports) into baseline/_ref/ (git-ignored, travels with gpurun).

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
    # recorded Gemini responses (the "recorded responses" the network call is stubbed with) next to their pages
    n_json = 0
    for name in sorted(os.listdir(os.path.join(REF, "pages"))):
        if name.endswith(".json"):
            shutil.copyfile(os.path.join(REF, "pages", name), os.path.join(DST, "pages", name)); n_json += 1
    # the reference's own call site (backend/app: config + pipeline/pdf_extract + utils), UNMODIFIED, into the git-ignored
    # baseline/_ref — tests/test_gpu_callsite.py imports it from there with stub pdf2image / google.generativeai modules
    root = os.path.dirname(os.path.dirname(HERE))
    app_src, app_dst = "/root/reference/backend/app", os.path.join(root, "baseline", "_ref", "backend", "app")
    for rel in ("__init__.py", "config.py", "pipeline/__init__.py", "pipeline/pdf_extract.py", "pipeline/utils.py"):
        os.makedirs(os.path.dirname(os.path.join(app_dst, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(app_src, rel), os.path.join(app_dst, rel))
    print(f"copied {n_json} recorded responses and the reference call site to {app_dst}")


if __name__ == "__main__":
    main()
