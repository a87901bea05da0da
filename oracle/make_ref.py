"""TEST INFRASTRUCTURE -- stages the reference's OWN hot-path modules under oracle/_ref/ so that they travel to the GPU box.

    python -m oracle.make_ref            (also run by __graft_entry__.build() whenever /root/reference is present)

The reference is pure Python; its path needs exactly four of its source files plus package markers:

    src/probabilistic_inference/probabilistic_inference.py     build_predictor, RetinaNetProbabilisticPredictor
    src/probabilistic_inference/inference_utils.py             NMS / fusion / covariance / post-processing helpers
    src/probabilistic_modeling/probabilistic_retinanet.py      ProbabilisticRetinaNet + head
    src/probabilistic_modeling/modeling_utils.py               covariance_output_to_cholesky

They are copied BYTE FOR BYTE (sha256 recorded in oracle/_ref/MANIFEST.json) into oracle/_ref/src/, a directory that is
git-ignored -- no reference source ever enters this repository's history -- but not gpurun-ignored, so the GPU box
(which has no /root/reference) can run the unmodified reference as the CPU arm of bench.py (`--impl reference`,
cpu_baseline.kind == "reference") and re-run tests/test_oracle_live_reference.py.  The un-vendored third-party
packages the reference imports (detectron2, fvcore; not installable offline) are provided by this repository's
stand-in oracle/ref_shim/, which is first-party test code and is used in place, not copied.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "src/__init__.py",
    "src/probabilistic_inference/__init__.py",
    "src/probabilistic_inference/probabilistic_inference.py",
    "src/probabilistic_inference/inference_utils.py",
    "src/probabilistic_modeling/__init__.py",
    "src/probabilistic_modeling/probabilistic_retinanet.py",
    "src/probabilistic_modeling/modeling_utils.py",
    "LICENSE.md",
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def staged():
    return os.path.isfile(os.path.join(DEST, "src", "probabilistic_inference", "probabilistic_inference.py"))


def stage(reference_root="/root/reference", verbose=False):
    """Copy the files (idempotent). Returns DEST, or None when the reference tree is absent."""
    if not os.path.isdir(os.path.join(reference_root, "src", "probabilistic_inference")):
        return DEST if staged() else None
    manifest = {}
    for rel in FILES:
        src = os.path.join(reference_root, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "sha256": manifest,
                   "note": "byte-for-byte copies of the reference's own files; git-ignored test infrastructure"}, f, indent=1)
    if verbose:
        print("staged %d reference files under %s" % (len(manifest), DEST))
    return DEST


if __name__ == "__main__":
    out = stage(verbose=True)
    if out is None:
        sys.exit("reference tree not available and nothing staged")
