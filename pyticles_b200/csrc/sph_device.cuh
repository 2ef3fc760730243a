// pyticles_b200 -- device helpers shared by the kernel translation units (sph_kernels.cu, sph_tiles.cu):
// 256-bit row loads, the reference's minimum image and pair predicate, and the block-Morton cell codes.
#ifndef SPH_DEVICE_CUH
#define SPH_DEVICE_CUH

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "pyticles_b200.h"

#define SPH_PI 3.14159265358979323846

namespace {

constexpr int kBlock = 256;
constexpr int kNlWarps = 8;          // warps per block in the neighbour pass
constexpr int kNlWin = 512;          // candidates staged per warp per window

struct __align__(16) d2 { double x, y; };

// One 256-bit load (LDG.E.ENL2.256) of a 32-byte-aligned row of four doubles: a gathered row
// costs one L1 request instead of two 128-bit ones -- the L1 data pipe is what bounds the
// density and force passes (profiles/r1a_kernels.txt).
__device__ __forceinline__ void load4(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

__device__ __forceinline__ void store4(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __forceinline__ uint32_t pdep32(uint32_t v, uint32_t mask)
{
    uint32_t r = 0;
    while (mask) {
        const uint32_t low = mask & (0u - mask);
        if (v & 1u) r |= low;
        v >>= 1;
        mask ^= low;
    }
    return r;
}

__device__ __forceinline__ uint32_t pext32(uint32_t v, uint32_t mask)
{
    uint32_t r = 0, bit = 1;
    while (mask) {
        const uint32_t low = mask & (0u - mask);
        if (v & low) r |= bit;
        bit <<= 1;
        mask ^= low;
    }
    return r;
}

// neighbour_list.py:111-122 -- one shift, strict comparisons against L/2.
__device__ __forceinline__ double min_image(double d, double L, double half)
{
    if (d > half) d = __dsub_rn(d, L);
    if (d < -half) d = __dadd_rn(d, L);
    return d;
}

// rsq exactly as numpy forms it: (dx*dx + dy*dy) + dz*dz, every operation rounded.
__device__ __forceinline__ double rsq_exact(double dx, double dy, double dz)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// The reference's pair predicate on the reference's operands (neighbour_list.py:170-178).
__device__ __forceinline__ bool pair_exact(const sph_grid &g, const double *pos4, int a, int j)
{
    double ax, ay, az, am, bx, by, bz, bm;
    load4(pos4 + 4 * (size_t)a, ax, ay, az, am);
    load4(pos4 + 4 * (size_t)j, bx, by, bz, bm);
    const double dx = min_image(__dsub_rn(bx, ax), g.box[0], g.box[0] / 2.);
    const double dy = min_image(__dsub_rn(by, ay), g.box[1], g.box[1] / 2.);
    const double dz = min_image(__dsub_rn(bz, az), g.box[2], g.box[2] / 2.);
    return rsq_exact(dx, dy, dz) < g.thr;
}

// ------------------------------------------------------------------ binning
struct CellLoc {
    uint32_t code;
    float rx, ry, rz;
    uint32_t flags;
    int cx;              // local x cell layer (the slab decomposition's boundary layers are picked by it)
    bool interior;       // every dimension: >= 5 layers and not the first or last GLOBAL layer
};

// Cell code = (block index << lbits) | Morton code inside the block.  Blocks are 2^lb x 2^lb x 2^lb
// cells (lb <= 3) numbered row-major, so the code space exceeds the real cell count by a few per
// cent only, whatever the number of layers; inside a block the order is a (generalised) Morton curve.
__device__ __forceinline__ uint32_t cell_code(const sph_grid &g, int cx, int cy, int cz)
{
    const uint32_t lx = g.lb[0], ly = g.lb[1], lz = g.lb[2];
    const uint32_t blk = (((uint32_t)cz >> lz) * g.nblk[1] + ((uint32_t)cy >> ly)) * g.nblk[0] + ((uint32_t)cx >> lx);
    const uint32_t loc = pdep32((uint32_t)cx & ((1u << lx) - 1u), g.mask[0]) |
                         pdep32((uint32_t)cy & ((1u << ly) - 1u), g.mask[1]) |
                         pdep32((uint32_t)cz & ((1u << lz) - 1u), g.mask[2]);
    return (blk << g.lbits) | loc;
}

// blk -> (bx, by, bz) with multiply-high division: magic = ceil(2^32 / nblk) is exact here because
// blk * nblk < 2^31 (sph_grid_plan rejects larger code spaces).
__device__ __forceinline__ void block_coords(const sph_grid &g, uint32_t blk, uint32_t &bx, uint32_t &by, uint32_t &bz)
{
    const uint32_t t = g.nblk[0] == 1u ? blk : __umulhi(blk, g.magic0);
    bx = blk - t * g.nblk[0];
    bz = g.nblk[1] == 1u ? t : __umulhi(t, g.magic1);
    by = t - bz * g.nblk[1];
}

__device__ __forceinline__ void cell_coords(const sph_grid &g, uint32_t code, int &cx, int &cy, int &cz)
{
    uint32_t bx, by, bz;
    block_coords(g, code >> g.lbits, bx, by, bz);
    cx = (int)((bx << g.lb[0]) | pext32(code, g.mask[0]));
    cy = (int)((by << g.lb[1]) | pext32(code, g.mask[1]));
    cz = (int)((bz << g.lb[2]) | pext32(code, g.mask[2]));
}

__device__ __forceinline__ int cell_coord(const sph_grid &g, int d, double x, float &rel, uint32_t &flags)
{
    const double L = g.box[d];
    if (!(x >= 0.0 && x < L)) {
        flags |= SPH_F_OUT_OF_BOX;
        if (!(x >= -0.25 * L && x <= 1.25 * L)) flags |= SPH_F_OUT_OF_RANGE;
        if (!(fabs(x) <= 1.0e300)) flags |= SPH_F_NONFINITE;
    }
    double f = floor(x * g.inv_w[d]);
    if (!(fabs(f) < 4.0e15)) f = 0.0;                 // NaN / absurd: any cell, exact path decides
    const long long cu = (long long)f;
    rel = (float)(x - (double)cu * g.w[d]);
    long long cg = cu % g.nc[d];
    if (cg < 0) cg += g.nc[d];
    int cl = (int)cg - g.lo[d];
    if (cl < 0) cl += g.nc[d];
    if (cl >= g.ncl[d]) { flags |= SPH_F_OUT_OF_SLAB; cl = g.ncl[d] - 1; }
    return cl;
}

__device__ __forceinline__ CellLoc locate(const sph_grid &g, double x, double y, double z)
{
    CellLoc c;
    c.flags = 0;
    const int cc[3] = {cell_coord(g, 0, x, c.rx, c.flags), cell_coord(g, 1, y, c.ry, c.flags),
                       cell_coord(g, 2, z, c.rz, c.flags)};
    c.code = cell_code(g, cc[0], cc[1], cc[2]);
    c.cx = cc[0];
    c.interior = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int cg = cc[d] + g.lo[d];
        if (cg >= g.nc[d]) cg -= g.nc[d];
        c.interior = c.interior && g.nc[d] >= 5 && cg >= 1 && cg <= g.nc[d] - 2;
    }
    return c;
}

__device__ __forceinline__ double lucy_norm3(double h)
{
    return 105. / (SPH_PI * 16. * (h * h * h));       // spkernel.py:99
}

}  // namespace

#endif  // SPH_DEVICE_CUH
