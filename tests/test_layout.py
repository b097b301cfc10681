"""CPU tier: the compile-time storage layout (csrc/trajectory.h `Slots`) of every kernel
instantiation respects the hardware budgets it is built around.  A small host program prints the
layout constants; nothing here needs a GPU."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#define TRPL_HOST_EMU 1
#include <stdio.h>
#include "%(root)s/metrotrpl_b200/csrc/trajectory.h"
using namespace trpl;
template <int NPL, int MODEL> void show() {
  typedef Slots<NPL, MODEL> S;
  printf("%%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d\n", NPL, MODEL, S::KSTRIDE, S::KP, S::FACP,
         S::TM_COUNT, S::TM_COLS, S::COUNT, S::BYTES, S::KCAP, (int)S::K_IN_TM, (int)S::FAC_IN_TM, S::PM_TM_PAIRS,
         S::PM_SM_PAIRS, S::K_TM_STAGES, S::TM_FAC);
}
int main() { show<1,0>(); show<2,0>(); show<4,0>(); show<8,0>(); show<1,1>(); show<2,1>(); show<4,1>(); show<8,1>(); }
'''


def layouts(team=False):
    """Layout rows of the one-warp instantiations, or (team) of the same source compiled with the
    two-warp vocabulary (csrc/team_kernels.cu: only 4 nodes per lane is instantiated there)."""
    tmp = tempfile.mkdtemp()
    src = os.path.join(tmp, "slots.cpp")
    with open(src, "w") as f:
        f.write(PROGRAM % {"root": ROOT})
    exe = os.path.join(tmp, "slots")
    subprocess.check_call(["g++", "-std=c++17"] + ([f"-DTRPL_TEAM={int(team)}"] if team else []) + ["-o", exe, src])
    rows = []
    for line in subprocess.check_output([exe], text=True).strip().splitlines():
        v = [int(x) for x in line.split()]
        rows.append(dict(zip(["npl", "model", "kstride", "kp", "facp", "tm_count", "tm_cols", "sm_pairs",
                              "sm_bytes", "kcap", "k_in_tm", "fac_in_tm", "pm_tm", "pm_sm", "k_tm_stages",
                              "tm_fac"], v)))
    return rows


def test_every_instantiation_fits_the_sm():
    rows = layouts()
    assert len(rows) == 8
    for r in rows:
        # tensor memory: one CTA of four warps allocates tm_cols columns (a power of two >= 32 that
        # holds the slice); CTAs resident on an SM share 512 columns
        assert r["tm_cols"] in (32, 64, 128, 256, 512) and 4 * r["tm_count"] <= r["tm_cols"]
        ctas_tm = 512 // r["tm_cols"]
        assert ctas_tm >= 1
        # shared memory: at least one warp's slice fits the 227 KB a CTA may use
        assert r["sm_bytes"] == r["sm_pairs"] * 512 and r["sm_bytes"] <= 227 * 1024
        # the explicit Runge-Kutta path keeps seven stages from KBASE on
        assert r["kcap"] >= 7 * r["kstride"]
        # every run that moves in .x16 groups starts on a four-pair boundary
        assert r["tm_fac"] % 4 == 0
        # the PCR multipliers are either all in registers or split whole levels tensor/shared
        assert r["pm_tm"] % 4 == 0 and (r["pm_tm"] + r["pm_sm"] in (0, 22, 24))


def test_headline_instantiation_layout():
    """nx = 128, 'std': factors (32 pairs) + multipliers (24) + K1, K2 (8) fill the 64-pair tensor
    memory budget of two CTAs per SM; K3..K5 region, exchange pairs and coefficients in shared memory."""
    r = [x for x in layouts() if x["npl"] == 4 and x["model"] == 0][0]
    assert r["fac_in_tm"] == 1 and r["k_in_tm"] == 0
    assert r["pm_tm"] == 24 and r["pm_sm"] == 0 and r["k_tm_stages"] == 2
    assert r["tm_count"] == 64 and r["tm_cols"] == 256
    assert 8 * r["sm_bytes"] <= 227 * 1024          # eight trajectories per SM


def test_two_warp_team_layout():
    """nx = 129..256 (csrc/team_kernels.cu): each of the two warps holds 4 nodes per lane, so a warp's
    tensor-memory slice is the headline kernel's 64 pairs (two CTAs = four trajectories per SM); the
    reduced system has 64 rows, hence 6 PCR levels of multipliers (26 pairs + padding); a TEAM's
    shared-memory slice is 64 lanes wide and four of them, plus the 8 KB mailbox, fit an SM."""
    for model in (0, 1):
        r = [x for x in layouts(team=2) if x["npl"] == 4 and x["model"] == model][0]
        assert r["fac_in_tm"] == 1
        assert r["tm_count"] <= 64 and r["tm_cols"] == 256            # two CTAs per SM
        assert r["pm_tm"] + r["pm_sm"] in (26, 28) and r["pm_tm"] % 4 == 0
        assert r["sm_bytes"] == r["sm_pairs"] * 1024                    # 64 lanes x 16 B per pair
        assert r["kcap"] >= 7 * r["kstride"]
        # two CTAs per SM, two teams per CTA, 8 KB of static mailbox + 1 KB reserved per CTA
        assert 2 * (2 * r["sm_bytes"] + 8208 + 1024) <= 228 * 1024


def test_four_warp_team_layout():
    """nx = 257..512 (csrc/team4_kernels.cu): 128 lanes, 7 PCR levels (30 pairs of multipliers + padding),
    one trajectory per CTA; two CTAs with their 12 KB mailboxes fit an SM."""
    for model in (0, 1):
        r = [x for x in layouts(team=4) if x["npl"] == 4 and x["model"] == model][0]
        assert r["fac_in_tm"] == 1
        assert r["tm_count"] <= 64 and r["tm_cols"] == 256
        assert r["pm_tm"] + r["pm_sm"] in (30, 32) and r["pm_tm"] % 4 == 0
        assert r["sm_bytes"] == r["sm_pairs"] * 2048                    # 128 lanes x 16 B per pair
        assert r["kcap"] >= 7 * r["kstride"]
        assert 2 * (r["sm_bytes"] + 12304 + 1024) <= 228 * 1024

