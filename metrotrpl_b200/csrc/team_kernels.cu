// team_kernels.cu - the trajectory kernels for grids of 129..256 nodes: a TEAM of two warps per
// trajectory (64 lanes x 4 nodes) instead of one warp with 8 nodes per lane.  (team4_kernels.cu: the
// same for 257..512 nodes with four warps per trajectory.)
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
#define TRPL_TEAM_NAME(x) trpl_team_##x
#include "team_kernels.inc"
