"""Stage the UNMODIFIED reference scripts of the hot path under oracle/_ref/ (git-ignored, NOT gpurun-ignored) so that
the literal reference classes can run on the GPU box, where /root/reference does not exist:

    python oracle/fetch_ref.py          # build container only; __graft_entry__.build() runs it when /root/reference is present

TEST / MEASUREMENT INFRASTRUCTURE -- not product code, never tracked: the files are byte copies made by this recipe from
where they lie under /root/reference, into a directory listed in .gitignore (the repo holds no reference source).  Users:
`bench.py --impl reference` and its `cpu_baseline` / `eager_b200` legs (the reference's own DLADMMNet on the host cores and,
unmodified, on the B200 through stock PyTorch eager).  oracle/load_reference.py falls back to this directory when
/root/reference is absent; with neither, those legs time the oracle port and say so (`kind: "port"`).
"""
import os
import shutil
import sys

SRC = os.environ.get("DLADMM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# the files that declare the hot path (+ the evaluation classes and their mu updaters, 8(f)-1/2)
FILES = ["main_syn_l1l1_scalar.py", "main_syn_l1l1_full.py", "main_syn_l1l1_scalar_tied.py", "main_syn_lasso_scalar.py",
         "main_lena.py", "main_syn_l1l1_ltheta.py", "gen_syn_data.py", "test_syn_l1l1_scalar.py",
         "test_syn_l1l1_newS_Acols.py", "mu_updater.py", "main_syn_scalar_newS_layerwise.py"]


def fetch(verbose=True):
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        if verbose:
            print("reference tree not found at %s: nothing staged" % SRC)
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    with open(os.path.join(DST, "README"), "w") as fh:
        fh.write("byte copies of %d files of xhchrn/D-LADMM staged by oracle/fetch_ref.py; git-ignored\n" % len(FILES))
    if verbose:
        print("staged %d reference files under %s" % (len(FILES), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)
