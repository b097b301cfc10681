// team4_kernels.cu - the trajectory kernels for grids of 257..512 nodes: a team of FOUR warps per
// trajectory (128 lanes x 4 nodes), one trajectory per CTA.  Same source, same vocabulary idea as
// team_kernels.cu (see there and simt.h): 7 PCR levels, three warp boundaries per mailbox slot, four
// partials per reduction, 128-bit lane masks, named barrier of 128 threads.
#define TRPL_TEAM 4
#define trpl trpl_team4
#define simt simt_team4
#define TRPL_TEAM_NAME(x) trpl_team4_##x
#include "team_kernels.inc"
