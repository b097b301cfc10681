"""Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, copied from
/root/reference so that they travel to the GPU box (oracle/_ref/ is git-ignored, not
gpurun-ignored; nothing is copied into the tracked tree).

    python oracle/make_ref.py            # run in the build container; __graft_entry__.build() calls it

bench.py --impl reference and bench.py's cpu_baseline leg drive these files' own
trial_move_evaluation.eval_trial_move (numba right-hand side + SciPy LSODA + likelihood) when they
are present (`kind: "reference"`), and fall back to the pinned oracle port (`kind: "port"`) when not.
"""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# eval_trial_move and everything it imports (trial_move_evaluation.py:4-7)
FILES = ["trial_move_evaluation.py", "forward_solver.py", "utils.py", "laplace.py", "sim_utils.py",
         "trial_move_generation.py"]


def make_ref(verbose=False):
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
    if verbose:
        print(f"copied {len(FILES)} reference modules to {DST}")
    return True


def available():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


if __name__ == "__main__":
    sys.exit(0 if make_ref(verbose=True) else 1)
