"""Stage the reference's own hot-path Python for the GPU box (recipe for `oracle/_ref/`).

TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product package.

shrebox/B-cosification is pure Python on ATen: there is nothing to compile.  To let `bench.py --impl reference` and the
`cpu_baseline` leg time THE REFERENCE'S OWN CODE on the GPU box's host cores (where /root/reference does not exist), this
recipe packs the unmodified hot-path modules, read where they lie under /root/reference, into ONE build artefact:

    oracle/_ref/bcos_reference.zip      (git-ignored like every built artefact; travels with the gpurun snapshot)
    oracle/_ref/MANIFEST.json           (path -> sha256 of each packed file, reference root, file count)

`oracle/refload.py` imports from the archive (zipimport) when /root/reference is absent.  Nothing is copied into the tracked
tree.  Run by `__graft_entry__.build()` whenever /root/reference is present, or by hand:

    python oracle/stage_ref.py
"""
import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "bcos_reference.zip")
MANIFEST = os.path.join(OUT_DIR, "MANIFEST.json")
REF = os.environ.get("BCOS_REFERENCE_ROOT", "/root/reference")

# the forward / explanation path and the model builders that consume it; package __init__ files are
# NOT packed -- refload seeds empty namespace packages because the real ones import training-only dependencies
PACK = [
    ("bcos/modules", lambda f: f.endswith(".py")),
    ("bcos/models", lambda f: f in ("resnet.py", "densenet.py", "vit.py", "standard_models.py")),
    ("bcos", lambda f: f in ("common.py", "version.py")),
    ("CLIP/clip", lambda f: f == "model.py"),
    ("", lambda f: f in ("bcosify.py", "bcosify_vit.py")),
]


def stage(ref_root: str = REF, quiet: bool = False) -> str:
    if not os.path.isdir(os.path.join(ref_root, "bcos", "modules")):
        raise RuntimeError(f"reference checkout not found at {ref_root}")
    os.makedirs(OUT_DIR, exist_ok=True)
    files = []
    for rel_dir, keep in PACK:
        d = os.path.join(ref_root, rel_dir)
        for f in sorted(os.listdir(d)):
            if os.path.isfile(os.path.join(d, f)) and keep(f):
                files.append(os.path.join(rel_dir, f) if rel_dir else f)
    for base, _dirs, names in sorted(os.walk(os.path.join(ref_root, "bcos", "modules"))):      # sub-packages (norms/)
        rel_dir = os.path.relpath(base, ref_root)
        for f in sorted(names):
            rel = os.path.join(rel_dir, f)
            if f.endswith(".py") and rel not in files:
                files.append(rel)
    manifest = {}
    tmp = ARCHIVE + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for rel in files:
            with open(os.path.join(ref_root, rel), "rb") as fh:
                data = fh.read()
            manifest[rel] = hashlib.sha256(data).hexdigest()
            info = zipfile.ZipInfo(rel, date_time=(1980, 1, 1, 0, 0, 0))      # reproducible archive
            info.compress_type = zipfile.ZIP_DEFLATED
            z.writestr(info, data)
    os.replace(tmp, ARCHIVE)
    with open(MANIFEST, "w") as fh:
        json.dump({"reference_root": ref_root, "files": manifest, "count": len(manifest)}, fh, indent=1, sort_keys=True)
    if not quiet:
        print(f"staged {len(manifest)} reference files into {ARCHIVE} ({os.path.getsize(ARCHIVE) / 1e3:.1f} kB)")
    return ARCHIVE


if __name__ == "__main__":
    stage()
    sys.exit(0)
