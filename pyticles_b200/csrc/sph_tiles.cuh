// pyticles_b200 -- internal interface between the C ABI (sph_kernels.cu) and the cell-group
// ("tile") neighbour pass (sph_tiles.cu).  Not part of the public boundary.
#ifndef SPH_TILES_CUH
#define SPH_TILES_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "pyticles_b200.h"

namespace sph_tiles {

// Can this grid / buffer set use the tile kernel at all?  (>= 3 cell layers in every dimension, not
// disabled by SPH_TILES=0.)  Data-dependent limits (window, cell and list capacities, positions far
// outside the box) are found on the device and raise SPH_F_TILE_FALLBACK in the status block; the
// general kernel then redoes the pass.
bool eligible(const sph_grid *g, const sph_buffers *b);

// Neighbour pass over cell groups: fills the warp-transposed ELL rows and counts of sph_buffers.
int launch_list(const sph_grid *g, const sph_buffers *b, cudaStream_t s);

// sph_buffers.group_tab: 16 words per group of 8 cell codes (12 window-layer code contributions, 3 base cell
// coordinates, 1 spare); geometry only.
bool grid_has_groups(const sph_grid *g);
int64_t group_tab_elems(const sph_grid *g);
int fill_group_table(const sph_grid *g, uint32_t *tab, cudaStream_t s);

}  // namespace sph_tiles

// Tensor-core pre-filter variant of the same pass (sph_tiles_mma.cu; SPH_TILES=2 selects it): same contract.
namespace sph_tiles_mma {
bool eligible(const sph_grid *g, const sph_buffers *b);
int launch_list(const sph_grid *g, const sph_buffers *b, cudaStream_t s);
}  // namespace sph_tiles_mma

#endif  // SPH_TILES_CUH
