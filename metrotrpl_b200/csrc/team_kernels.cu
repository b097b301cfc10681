// team_kernels.cu - the trajectory kernels for grids of 129..256 nodes: a TEAM of two warps per
// trajectory (64 lanes x 4 nodes) instead of one warp with 8 nodes per lane.
//
// Why: with 8 nodes per lane the five state-sized vectors of a RODAS4 stage (24 doubles each with
// the trap occupancy) do not fit the register file, and the factor blocks of one trajectory fill a
// warp's share of tensor memory at ONE CTA per SM - round 2's ncu of that instantiation: 9% of the
// instructions were spill loads/stores, 629 MB of spill lines per launch, one warp per scheduler,
// 0.10 of the FP64 peak (profiles/r02_ncu_traps_nx256_kernel.md).  Two warps with 4 nodes per lane
// hold the same trajectory with the register and tensor-memory footprint of the headline kernel:
// no spills, two CTAs (four trajectories, eight warps) per SM.
//
// How: this unit compiles the SAME integrator source (trajectory.h, model.h, blocktri.h, irf.h,
// explicit.h) against the two-warp vocabulary of simt.h (TRPL_TEAM == 2): lane_id() is 0..63, the
// reduced system has 64 rows (6 PCR levels), neighbour shuffles hand one value across the warp
// boundary through a shared-memory mailbox, reductions combine two partials there, warp_sync() is
// a named barrier of the team (bar.sync 1+team, 64).  Tensor memory stays lane-private per warp.
// The namespaces are renamed so that nothing here collides with the one-warp unit's templates.
#define TRPL_TEAM 2
#define trpl trpl_team
#define simt simt_team
#include <cuda_runtime.h>
#include <algorithm>
#include <string.h>

#include "../../include/metrotrpl_b200.h"
#include "trajectory.h"
#include "explicit.h"
#include "kernel_common.h"

namespace {

constexpr int TEAMS_PER_CTA = WARPS_PER_CTA / 2;

// next trajectory of the work queue, claimed by the team's lane 0 (same queue order as trpl_kernels.cu)
__device__ __forceinline__ int claim(const KernelArgs& a, int* counter, int limit) {
  int q = 0;
  if (lane_id() == 0) q = atomicAdd(counter, 1);
  q = (int)lane0((double)q);
  return q < limit ? q : -1;
}

template <int NPL, int MODEL, bool FULL>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA, CTAS_PER_SM) trpl_team_forward_kernel(const KernelArgs a) {
  typedef Slots<NPL, MODEL> SL;
  extern __shared__ double2 smem[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int team = warp >> 1;
  TrajMem mem{LaneMem{smem + team * (SL::COUNT * LANES)}, tmem_acquire<SL>(warp)};
  const bool allow_defer = a.defer_list != nullptr;
  for (;;) {
    int traj = claim(a, a.counter, a.n_traj);
    if (traj < 0) break;
    if (a.queue) {
      traj = a.queue[traj];
    } else {
      const int n_sets_q = a.n_traj / a.n_meas;
      const int qm = traj / n_sets_q;
      traj = (traj - qm * n_sets_q) * a.n_meas + a.meas_order[qm];
    }
    TrajIn in;
    setup_traj(a, traj, team, in);
    TrajOut out;
    TrajMid mid;
    if (run_trajectory<NPL, MODEL, FULL>(in, a.opt, mem, out, mid, allow_defer)) {
      if (lane_id() == 0) a.defer_list[atomicAdd(a.defer_count, 1)] = traj;     // non-stiff: explicit path
      continue;
    }
    finish_traj(a, traj, team, in, mid, out);
  }
  tmem_release<SL>(warp, mem.tm);
}

template <int NPL, int MODEL, bool FULL>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA, CTAS_PER_SM) trpl_team_explicit_kernel(const KernelArgs a) {
  typedef Slots<NPL, MODEL> SL;
  extern __shared__ double2 smem[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int team = warp >> 1;
  TrajMem mem{LaneMem{smem + team * (SL::COUNT * LANES)}, tmem_acquire<SL>(warp)};
  const int n = *a.defer_count;
  for (;;) {
    const int q = claim(a, a.counter + 1, n);
    if (q < 0) break;
    const int traj = a.defer_list[q];
    TrajIn in;
    setup_traj(a, traj, team, in);
    TrajOut out;
    TrajMid mid;
    run_trajectory_explicit<NPL, MODEL, FULL>(in, a.opt, mem, out, mid);
    out.status |= ST_EXPLICIT;
    finish_traj(a, traj, team, in, mid, out);
  }
  tmem_release<SL>(warp, mem.tm);
}

// every call site of a team primitive owns one mailbox slot (simt.h)
static_assert(__COUNTER__ <= simt::TEAM_SLOTS, "more team-primitive call sites than mailbox slots");

template <int MODEL, bool FULL>
cudaError_t run(const KernelArgs& a, int grid, size_t smem, cudaStream_t stream, bool explicit_pass) {
  auto kern = explicit_pass ? trpl_team_explicit_kernel<4, MODEL, FULL> : trpl_team_forward_kernel<4, MODEL, FULL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, 32 * WARPS_PER_CTA, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace

// ---- what trpl_kernels.cu calls (not part of the C ABI: hidden symbols) ----
extern "C" {

// shared-memory bytes and tensor-memory columns of one CTA (two teams), trajectories per CTA
__attribute__((visibility("hidden"))) void trpl_team_plan(int model, size_t* smem, int* tm_cols, int* teams_per_cta) {
  if (model == TRPL_MODEL_TRAPS) {
    typedef Slots<4, MODEL_TRAPS> SL;
    *smem = (size_t)TEAMS_PER_CTA * SL::BYTES; *tm_cols = SL::TM_COUNT > 0 ? TmCta<SL>::COLS : 0;
  } else {
    typedef Slots<4, MODEL_STD> SL;
    *smem = (size_t)TEAMS_PER_CTA * SL::BYTES; *tm_cols = SL::TM_COUNT > 0 ? TmCta<SL>::COLS : 0;
  }
  *teams_per_cta = TEAMS_PER_CTA;
}

// `args` points at a KernelArgs (kernel_common.h: the same struct in both units)
__attribute__((visibility("hidden"))) int trpl_team_run(int model, int full, const void* args, size_t args_size,
                                                         int grid, size_t smem, void* stream, int explicit_pass) {
  if (args_size != sizeof(KernelArgs)) return (int)cudaErrorInvalidValue;
  KernelArgs a;
  memcpy(&a, args, sizeof(a));
  cudaStream_t st = (cudaStream_t)stream;
  const bool ex = explicit_pass != 0;
  if (model == TRPL_MODEL_TRAPS)
    return (int)(full ? run<MODEL_TRAPS, true>(a, grid, smem, st, ex) : run<MODEL_TRAPS, false>(a, grid, smem, st, ex));
  return (int)(full ? run<MODEL_STD, true>(a, grid, smem, st, ex) : run<MODEL_STD, false>(a, grid, smem, st, ex));
}

}  // extern "C"
