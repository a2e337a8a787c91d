"""TEST / BENCH INFRASTRUCTURE ONLY -- populate oracle/_ref/ with the UNMODIFIED reference package.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference exists)

The reference's hot path (isdf/modules/trainer.py:951-1016 and what it imports) is pure Python on top of torch, so
"building" it is a verbatim copy of the `isdf/` package directory from the read-only checkout into oracle/_ref/isdf
(git-ignored, NOT gpurun-ignored: it travels to the GPU box next to the built .so).  Nothing is edited; a manifest with
the SHA-256 of every copied file is written next to it so the copy can be audited against /root/reference.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import what lands here (through oracle/ref_shim.py).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ISDF_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def populate(verbose=False):
    src_pkg = os.path.join(SRC, "isdf")
    if not os.path.isdir(os.path.join(src_pkg, "modules")):
        return False                      # GPU box: the reference checkout is absent, the prebuilt copy is used
    dst_pkg = os.path.join(DST, "isdf")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    manifest = {}
    for root, dirs, files in os.walk(src_pkg):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, SRC)
        os.makedirs(os.path.join(DST, rel), exist_ok=True)
        for f in files:
            if not f.endswith((".py", ".json", ".txt", ".xml")):
                continue
            s, d = os.path.join(root, f), os.path.join(DST, rel, f)
            shutil.copyfile(s, d)
            manifest[os.path.join(rel, f)] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: %d files copied verbatim from %s" % (len(manifest), SRC))
    return True


if __name__ == "__main__":
    ok = populate(verbose=True)
    sys.exit(0 if ok else 1)
