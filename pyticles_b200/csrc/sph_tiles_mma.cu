// pyticles_b200 -- cell-group ("tile") neighbour pass for sm_100a, tensor-core pre-filter variant (SPH_TILES=2).
//
// EXPERIMENT of round 2, kept selectable because it is exact and tested, NOT the default: it replaces the scalar
// candidate test of sph_tiles.cu by HMMA tiles, which cuts the test instructions 4x, but writing the hits out of the
// fragment layout (two home particles and two tile rows per lane: a third of the lanes busy) costs what the tests
// saved -- 5.7 ms against 4.4 ms on the 256^3 box (profiles/r2_tile_mma_*.txt).
//
// One block works on a GROUP of 2 x 2 x 2 cells (eight consecutive block-Morton codes), one warp per cell.  The
// block stages the 4 x 4 x 4 cells around the group ONCE in shared memory (a particle row is read from L2 8x
// instead of 27x, with coalesced loads).
//
// The candidate test is a distance pre-filter on the tensor cores.  For a candidate c and a home particle p, in
// coordinates u of the group's frame scaled by 1 / cell width (|u| <= 2),
//     d(c, p) = |u_c|^2 + |u_p|^2 - thr' - 2 u_c . u_p  =  rsq' - thr'
// is ONE dot product of two 16-vectors once each coordinate is split into two halves (fp16 hi + lo, products
// exact in the fp32 accumulator) and the squared norms into three:
//     A row of c  [ xh  yh  zh   xh   yh   zh   xl  yl |  zl  n_h n_l n_ll  1    1    1    0 ]
//     B col of p  [-2xh -2yh -2zh -2xl -2yl -2zl -2xh -2yh | -2zh  1   1   1  k_h  k_l  k_ll  0 ]     k = |u_p|^2 - thr'
// so mma.sync.m16n8k16 (HMMA.16816.F32) evaluates 16 candidates x 8 home particles = 128 tests per instruction,
// its A fragment fetched with one ldmatrix.x4.  The sign bit of each of a lane's four results IS the hit: it is
// funnel-shifted into a bit mask; nothing else happens per test (r1's scalar loop: 14 instructions per 32 tests,
// here about 10 per 128).  The staged window is laid out so that the three cell rows (4 cells each) above, at and
// below a home cell in one z plane are one contiguous run of candidates: a pass is 3 runs of about 7 tiles.  The
// fourth cell of a row and a tile's overshoot into the next row lie two cells from the home cell: they cannot hit.
//
// Afterwards the masks are walked (find-first-set): a lane owns the hits of two home particles in the tile rows g and
// g + 8; the eight lanes that share a home particle concatenate their hits, offsets from a shuffle scan, into the
// particle's warp-transposed ELL row (the neighbour structure every other pass and the export use).
//
// Exactness is the one of the general kernel: d >= 0 rejects (rsq' >= thr' + band), d < -bw accepts, and a lane
// that saw |d| < bw -- the rigorous error band of the split arithmetic, tile_thresholds -- re-decides all its hits
// with the reference's fp64 predicate (pair_exact).  Cases outside the fixed capacities (> 64 particles in a cell,
// > 128 in the group, > 1152 in its window, > 16 tiles in a plane's run, positions far outside the box) raise
// SPH_F_TILE_FALLBACK and the general kernel redoes the pass.
//
// Tensor cores here are a pre-filter, not the roofline: the pass stays bound by instruction issue and HBM
// (DESIGN.md section 5); the legacy warp-level HMMA path is used because a tile is 16 x 8 (tcgen05 tiles start at
// 64 x 8 and would need the accumulators read back from tensor memory) and runs at 1.1 dense PFLOP/s on B200
// (tools/hmma_probe.cu), 50x what this pass asks of it.
//
// Reference semantics (file:line into the reference tree):
//   pair predicate     neighbour_list.py:105-123,170-178
#include <cuda_fp16.h>
#include <stdlib.h>

#include "sph_device.cuh"
#include "sph_tiles.cuh"

namespace {

constexpr int kTWarps = 8;           // warps per block = cells per group
constexpr int kTThreads = kTWarps * 32;
#ifndef SPH_TILE_BLOCKS
#define SPH_TILE_BLOCKS 4
#endif
constexpr int kTBlocks = SPH_TILE_BLOCKS;              // resident blocks per SM
constexpr int kTCap = kTBlocks >= 5 ? 1024 : 1152;    // staged candidates per group (64 cells, rows padded to 8)
constexpr int kTTail = 32;           // far-away dummy candidates behind the window (a run's last tile may overshoot)
constexpr int kTPart = 64;           // particles per cell the tile path handles
constexpr int kTHome = 128;          // particles in the eight home cells
constexpr int kTWords = 6;           // mask words per lane and pass: two (8 tiles each) per z plane

constexpr uint32_t kFull = 0xffffffffu;

struct TileArgs {
    int n, K;
    const uint32_t *cell_start;
    const float *rel4;
    const double *pos4;
    int32_t *nbr;
    int32_t *cnt;
    sph_status *status;
    float scale;             // 1 / largest cell width: staged coordinates are u = x * scale
    float thr_s;             // thr' = (thr + band) * scale^2
    float bw;                // hits with d >= -bw are settled in fp64
    const int32_t *perm;
    int n_owned;             // > 0 on a restricted grid: the first / last local x layer hold ghosts (no rows for them)
};

// shared memory of a block: [A | I32 | Bv | Head]
struct Head {
    uint32_t off[68];        // first staged candidate of window cell (wz*4 + wy)*4 + wx; rows of 4 cells start 8-aligned
    uint32_t start[64];      // first sorted particle of the window cell
    uint32_t cnt[64];
    uint32_t hslot[12];      // first home-vector slot of home cell (hz*2 + hy)*2 + hx; [8] = particles in the group
    uint32_t rowlen[16], rowbase[20];
    float shift[12];         // (i - 2) * w[d] at [4 d + i]: fp32 frame shift of window layer i
    uint32_t part[12];       // cell code contribution of window layer i of dimension d at [4 d + i]; ~0u: no such layer
    int gc[4];               // local cell coordinates of the group's base cell
};

constexpr size_t kBytesA = 32 * (size_t)(kTCap + kTTail);
constexpr size_t kBytesI = sizeof(uint32_t) * (size_t)(kTCap + kTTail);
constexpr size_t kBytesBv = 32 * (size_t)kTHome;
constexpr size_t kSmemList = kBytesA + kBytesI + kBytesBv + sizeof(Head);
static_assert(kTBlocks * (kSmemList + 1024) <= 227 * 1024, "blocks per SM");
static_assert(kBytesA % 16 == 0 && kBytesI % 16 == 0, "alignment");

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi)
{
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// x = hi + lo + (residual below 2^-21 |x|), both halves
__device__ __forceinline__ void split2(float x, float &hi, float &lo)
{
    hi = __half2float(__float2half_rn(x));
    lo = __half2float(__float2half_rn(x - hi));
}

__device__ __forceinline__ void split3(float x, float &a, float &b, float &c)
{
    a = __half2float(__float2half_rn(x));
    const float r = x - a;
    b = __half2float(__float2half_rn(r));
    c = __half2float(__float2half_rn(r - b));
}

// Row `row` of the A matrix: 32 bytes, its two 16-byte halves swapped in rows 4..7 (mod 8) so that the eight row
// addresses of an ldmatrix phase fall into distinct banks.
__device__ __forceinline__ void store_a_row(unsigned char *A, uint32_t row, uint4 lo, uint4 hi)
{
    uint4 *p = reinterpret_cast<uint4 *>(A + 32u * row);
    const uint32_t sw = (row >> 2) & 1u;
    p[sw] = lo;
    p[sw ^ 1u] = hi;
}

// ------------------------------------------------------------------ window of a group
// Fills Head (window cells, 8-aligned rows, frame shifts, base coordinates, home slots) and stages the window: the A
// row and the sorted index of every candidate, the B vector of every particle of the eight home cells, far-away dummy
// rows in the padding.  Returns false when a fixed capacity is exceeded (nothing usable was staged).
__device__ __forceinline__ bool tile_stage(const sph_grid &g, uint32_t c0, const TileArgs &a, Head *H, unsigned char *A,
                                           uint32_t *I32, uint32_t *Bv)
{
    const int t = threadIdx.x;
    if (t < 12) {
        // The cell code is additive over the dimensions: (block coordinate * block stride) << lbits plus the
        // in-block Morton bits.  Twelve threads work out the contribution of window layer i = t & 3 of
        // dimension d = t >> 2; the 64 cell codes are then three table look-ups and two adds each.
        const int d = t >> 2, i = t & 3;
        int cc[3];
        cell_coords(g, c0, cc[0], cc[1], cc[2]);
        if (t == 0) { H->gc[0] = cc[0]; H->gc[1] = cc[1]; H->gc[2] = cc[2]; }
        H->shift[t] = (float)((double)(i - 2) * g.w[d]);
        int c = cc[d] + i - 1;
        bool ok = true;
        if (c < 0) {
            if (g.wrap[d]) c += g.ncl[d]; else ok = false;
        } else if (c >= g.ncl[d]) {
            if (g.wrap[d]) c -= g.ncl[d]; else ok = false;
        }
        uint32_t part = ~0u;
        if (ok) {
            const uint32_t stride = d == 0 ? 1u : (d == 1 ? g.nblk[0] : g.nblk[0] * g.nblk[1]);
            part = ((((uint32_t)c >> g.lb[d]) * stride) << g.lbits) |
                   pdep32((uint32_t)c & ((1u << g.lb[d]) - 1u), g.mask[d]);
        }
        H->part[t] = part;
    }
    __syncthreads();
    if (t < 64) {
        const uint32_t px = H->part[t & 3], py = H->part[4 + ((t >> 2) & 3)], pz = H->part[8 + (t >> 4)];
        uint32_t st = 0, cn = 0;
        if (px != ~0u && py != ~0u && pz != ~0u) {
            const uint32_t code = px + py + pz;
            st = a.cell_start[code];
            cn = a.cell_start[code + 1] - st;
        }
        H->start[t] = st;
        H->cnt[t] = cn;
    }
    __syncthreads();
    if (t < 32) {
        // lanes 0..15: one window row (4 cells along x) each, padded to a multiple of 8 candidates
        uint32_t c[4] = {0, 0, 0, 0}, len = 0;
        if (t < 16) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { c[k] = H->cnt[4 * t + k]; len += c[k]; }
        }
        const uint32_t padded = (len + 7u) & ~7u;
        uint32_t inc = padded;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const uint32_t x = __shfl_up_sync(kFull, inc, o);
            if (t >= o) inc += x;
        }
        if (t < 16) {
            const uint32_t base = inc - padded;
            H->rowlen[t] = len;
            H->rowbase[t] = base;
            if (t == 15) H->rowbase[16] = inc;
            uint32_t o = base;
#pragma unroll
            for (int k = 0; k < 4; ++k) { H->off[4 * t + k] = o; o += c[k]; }
            if (t == 15) H->off[64] = inc;
        }
        // lanes 16..23: slots of the home vectors, home cell (hz, hy, hx) = window cell (1 + hz, 1 + hy, 1 + hx)
        const int hcell = t - 16;
        uint32_t hc = 0;
        if (hcell >= 0 && hcell < 8)
            hc = H->cnt[((1 + (hcell >> 2)) * 4 + (1 + ((hcell >> 1) & 1))) * 4 + (1 + (hcell & 1))];
        uint32_t hinc = hc;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const uint32_t x = __shfl_up_sync(kFull, hinc, o);
            if (hcell >= o) hinc += x;
        }
        if (hcell >= 0 && hcell < 8) {
            H->hslot[hcell] = hinc - hc;
            if (hcell == 7) H->hslot[8] = hinc;
        }
    }
    __syncthreads();
    const uint32_t total = H->rowbase[16];
    if (total > (uint32_t)kTCap || H->hslot[8] > (uint32_t)kTHome) return false;
    // dummy rows: the padding of every window row and the tail behind the window.  d = 60 + k > 0 against anything.
    {
        const uint4 lo = make_uint4(0u, 0u, 0u, 0u);
        const uint4 hi = make_uint4(pack_h2(0.f, 60.f), 0u, pack_h2(1.f, 1.f), pack_h2(1.f, 0.f));
        if (t < 128) {
            const int r = t >> 3, k = t & 7;
            const uint32_t at = H->rowbase[r] + H->rowlen[r] + (uint32_t)k;
            if (at < H->rowbase[r + 1]) { store_a_row(A, at, lo, hi); I32[at] = 0u; }
        } else if (t < 128 + kTTail) {
            const uint32_t at = total + (uint32_t)(t - 128);
            store_a_row(A, at, lo, hi);
            I32[at] = 0u;
        }
    }
    // warp w stages window cells 8w .. 8w+7, four cells per pass (8 lanes each)
    const int w = t >> 5, lane = t & 31;
    const float4 *rel = reinterpret_cast<const float4 *>(a.rel4);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int wc = w * 8 + pass * 4 + (lane >> 3);
        const uint32_t st = H->start[wc], cn = H->cnt[wc], dst = H->off[wc];
        const int wx = wc & 3, wy = (wc >> 2) & 3, wz = wc >> 4;
        const float fx = H->shift[wx], fy = H->shift[4 + wy], fz = H->shift[8 + wz];
        const bool home = wx >= 1 && wx <= 2 && wy >= 1 && wy <= 2 && wz >= 1 && wz <= 2;
        const uint32_t hs = home ? H->hslot[((wz - 1) * 2 + (wy - 1)) * 2 + (wx - 1)] : 0u;
        for (uint32_t k = lane & 7; k < cn; k += 8) {
            const float4 p = __ldg(rel + st + k);
            const float x = (p.x + fx) * a.scale, y = (p.y + fy) * a.scale, z = (p.z + fz) * a.scale;
            const float nn = fmaf(z, z, fmaf(y, y, x * x));
            float xh, xl, yh, yl, zh, zl, n0, n1, n2;
            split2(x, xh, xl);
            split2(y, yh, yl);
            split2(z, zh, zl);
            split3(nn, n0, n1, n2);
            const uint32_t one = pack_h2(1.f, 1.f);
            store_a_row(A, dst + k, make_uint4(pack_h2(xh, yh), pack_h2(zh, xh), pack_h2(yh, zh), pack_h2(xl, yl)),
                        make_uint4(pack_h2(zl, n0), pack_h2(n1, n2), one, pack_h2(1.f, 0.f)));
            I32[dst + k] = st + k;
            if (home) {
                float k0, k1, k2;
                split3(nn - a.thr_s, k0, k1, k2);
                uint4 *bv = reinterpret_cast<uint4 *>(Bv + 8u * (hs + k));
                bv[0] = make_uint4(pack_h2(-2.f * xh, -2.f * yh), pack_h2(-2.f * zh, -2.f * xl),
                                   pack_h2(-2.f * yl, -2.f * zl), pack_h2(-2.f * xh, -2.f * yh));
                bv[1] = make_uint4(pack_h2(-2.f * zh, 1.f), one, pack_h2(k0, k1), pack_h2(k2, 0.f));
            }
        }
    }
    __syncthreads();
    return true;
}

__device__ __forceinline__ void ldsm4(uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t &a3, uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
}

__device__ __forceinline__ void mma16816(float &d0, float &d1, float &d2, float &d3, uint32_t a0, uint32_t a1,
                                         uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                 "{%10,%11,%12,%13};"
                 : "=f"(d0), "=f"(d1), "=f"(d2), "=f"(d3)
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f), "f"(0.f), "f"(0.f), "f"(0.f));
}

// The candidates of a home cell: three runs, one per z plane, each the three window rows (4 cells each) around the
// cell, contiguous in the staged window.  A run is tested in tiles of 16; the masks of its first and second eight
// tiles are mask words 2 dz and 2 dz + 1 of the pass (a run of more than 16 tiles goes to the general kernel).
struct Run {
    uint32_t s;              // first candidate (window index, a multiple of 8)
    int nt;                  // tiles
};

__device__ __forceinline__ Run run_of(const Head *H, int hy, int hz, int dz)
{
    const int r0 = 16 * (hz + dz) + 4 * hy;                              // first cell of the three rows
    Run r;
    r.s = H->off[r0];
    r.nt = (int)((H->off[r0 + 12] - r.s + 15u) >> 4);
    return r;
}

// Bit b of mask word w: tile k of the word sits in nibble (tiles in the word - 1 - k); a nibble is (d0 d1 d2 d3):
// rows g, g, g + 8, g + 8 of the tile, home particles 2t, 2t + 1, 2t, 2t + 1.
struct Word {
    uint32_t base;           // window index of row g of the word's LAST tile (nibble 0)
};

__device__ __forceinline__ Word word_of(const Run &r, int half, int g)
{
    const int ntw = min(8, r.nt - 8 * half);
    Word x;
    x.base = r.s + 16u * (uint32_t)(8 * half + ntw - 1) + (uint32_t)g;
    return x;
}

__device__ __forceinline__ uint32_t widx_of(const Word &x, int b)
{
    return x.base - 16u * (uint32_t)(b >> 2) + ((b & 2) ? 0u : 8u);       // bits 3, 2: d0, d1 (row g); 1, 0: d2, d3 (g + 8)
}

// up to eight tiles from `addr` on: four sign bits per tile into the mask
__device__ __forceinline__ uint32_t test_tiles(uint32_t addr, int ntl, uint32_t b0, uint32_t b1, float &near)
{
    uint32_t mask = 0u;
#pragma unroll 2
    for (int k = 0; k < ntl; ++k, addr += 512u) {
        uint32_t a0, a1, a2, a3;
        float d0, d1, d2, d3;
        ldsm4(a0, a1, a2, a3, addr);
        mma16816(d0, d1, d2, d3, a0, a1, a2, a3, b0, b1);
        mask = __funnelshift_l(__float_as_uint(d0), mask, 1);
        mask = __funnelshift_l(__float_as_uint(d1), mask, 1);
        mask = __funnelshift_l(__float_as_uint(d2), mask, 1);
        mask = __funnelshift_l(__float_as_uint(d3), mask, 1);
        near = fminf(near, fminf(fminf(fabsf(d0), fabsf(d1)), fminf(fabsf(d2), fabsf(d3))));
    }
    return mask;
}

// ------------------------------------------------------------------ one cell of a staged group (one warp)
// Returns the longest row it wrote (0 when it gave up and raised SPH_F_TILE_FALLBACK).
__device__ __forceinline__ uint32_t tile_cell(const sph_grid &grid, const TileArgs &a, const Head *H,
                                              const unsigned char *A, const uint32_t *I32, const uint32_t *Bv)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int hx = w & 1, hy = (w >> 1) & 1, hz = w >> 2;
    const int wc = ((1 + hz) * 4 + (1 + hy)) * 4 + (1 + hx);
    // the last group of an odd layer count is half empty: its window cell is the periodic image of layer 0
    const bool exists = H->gc[0] + hx < grid.ncl[0] && H->gc[1] + hy < grid.ncl[1] && H->gc[2] + hz < grid.ncl[2];
    const int PC = exists ? (int)H->cnt[wc] : 0;
    if (PC == 0) return 0u;
    const uint32_t cs = H->start[wc], c0 = H->off[wc], hs = H->hslot[w];
    const Run run0 = run_of(H, hy, hz, 0), run1 = run_of(H, hy, hz, 1), run2 = run_of(H, hy, hz, 2);
    if (PC > kTPart || max(run0.nt, max(run1.nt, run2.nt)) > 16) {
        if (lane == 0) atomicOr(&a.status->flags, SPH_F_TILE_FALLBACK);
        return 0u;
    }
    if (a.n_owned > 0 && !grid.wrap[0]) {
        // slab decomposition: the first and the last local x layer hold the ghosts (sph_grid_restrict_x keeps one
        // ghost layer on each side of the owned ones); no rows are built for them
        const int cx = H->gc[0] + hx;
        if (cx == 0 || cx == grid.ncl[0] - 1) {
            for (int k = lane; k < PC; k += 32) a.cnt[cs + k] = 0;
            return 0u;
        }
    }
    const uint32_t a32 = smem_u32(A);
    // ldmatrix: lane L addresses row (L & 7) + 8 * ((L >> 3) & 1) of the tile, 16-byte half L >> 4 (swapped in rows
    // 4..7: run starts are multiples of 8, so the swap depends on the lane only)
    const uint32_t lrow = (uint32_t)(lane & 7) + (((uint32_t)lane >> 3) & 1u) * 8u;
    const uint32_t loff = a32 + lrow * 32u + ((((uint32_t)lane >> 4) ^ ((lrow >> 2) & 1u)) << 4);
    uint32_t wmax = 0;

#pragma unroll 1
    for (int p0 = 0; p0 < PC; p0 += 8) {
        const int P = min(8, PC - p0);
        // B fragment: words t and t + 4 of the home vector of particle g (a spare column repeats particle 0; masked below)
        const uint32_t hb = 8u * (hs + (uint32_t)p0 + (uint32_t)(g < P ? g : 0));
        const uint32_t b0 = Bv[hb + t], b1 = Bv[hb + t + 4];
        float near = INFINITY;
        uint32_t m[kTWords];                                             // static indices only: registers
        m[0] = test_tiles(loff + run0.s * 32u, min(8, run0.nt), b0, b1, near);
        m[1] = test_tiles(loff + run0.s * 32u + 4096u, run0.nt - 8, b0, b1, near);
        m[2] = test_tiles(loff + run1.s * 32u, min(8, run1.nt), b0, b1, near);
        m[3] = test_tiles(loff + run1.s * 32u + 4096u, run1.nt - 8, b0, b1, near);
        m[4] = test_tiles(loff + run2.s * 32u, min(8, run2.nt), b0, b1, near);
        m[5] = test_tiles(loff + run2.s * 32u + 4096u, run2.nt - 8, b0, b1, near);
        // drop the spare home columns and each home particle's test against itself, count
        const bool v0 = 2 * t < P, v1 = 2 * t + 1 < P;
        const uint32_t keep = (v0 ? 0xaaaaaaaau : 0u) | (v1 ? 0x55555555u : 0u);
        uint32_t self2 = 0u, self3 = 0u;                                 // bits to clear in words 2 and 3
#pragma unroll
        for (int sec = 0; sec < 2; ++sec) {
            // particle 2t + sec of the pass sits at window index c0 + p0 + 2t + sec, in the middle plane's run
            const uint32_t rel = c0 + (uint32_t)(p0 + 2 * t + sec) - run1.s;
            const int tile = (int)(rel >> 4), row = (int)(rel & 15u);
            if ((row & 7) == g) {
                const int half = tile >> 3, ntw = min(8, run1.nt - 8 * half);
                const uint32_t bit = 1u << (((ntw - 1 - (tile & 7)) * 4 + (3 - ((row >> 3) * 2 + sec))) & 31);
                if (half) self3 |= bit; else self2 |= bit;
            }
        }
        m[2] &= ~self2;
        m[3] &= ~self3;
        int cnt0 = 0, cnt1 = 0;
#pragma unroll
        for (int ws = 0; ws < kTWords; ++ws) {
            m[ws] &= keep;
            cnt0 += __popc(m[ws] & 0xaaaaaaaau);
            cnt1 += __popc(m[ws] & 0x55555555u);
        }
        const uint32_t as0 = cs + (uint32_t)(p0 + 2 * t), as1 = as0 + 1u;  // sorted indices of the two home particles
        // rare: some test of this lane fell into the error band -> the reference's fp64 predicate on all its hits
        if (near < a.bw && cnt0 + cnt1 > 0) {
            cnt0 = cnt1 = 0;
#pragma unroll
            for (int ws = 0; ws < kTWords; ++ws) {
                uint32_t mm = m[ws];
                const Run &r = ws < 2 ? run0 : (ws < 4 ? run1 : run2);
                const Word x = word_of(r, ws & 1, g);
                while (mm) {
                    const int b = 31 - __clz(mm);
                    mm ^= 1u << b;
                    if (!pair_exact(grid, a.pos4, (int)((b & 1) ? as0 : as1), (int)I32[widx_of(x, b)])) m[ws] ^= 1u << b;
                }
                cnt0 += __popc(m[ws] & 0xaaaaaaaau);
                cnt1 += __popc(m[ws] & 0x55555555u);
            }
        }
        // the eight lanes with the same t share two home particles: exclusive scan over g in a fixed order
        int inc0 = cnt0, inc1 = cnt1;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            const int x0 = __shfl_up_sync(kFull, inc0, o), x1 = __shfl_up_sync(kFull, inc1, o);
            if (lane >= o) { inc0 += x0; inc1 += x1; }
        }
        const int tot0 = __shfl_sync(kFull, inc0, 28 + t), tot1 = __shfl_sync(kFull, inc1, 28 + t);
        int32_t *e0p = a.nbr + ((size_t)(as0 >> 5) * (size_t)a.K + (size_t)(inc0 - cnt0)) * 32 + (as0 & 31);
        int32_t *e1p = a.nbr + ((size_t)(as1 >> 5) * (size_t)a.K + (size_t)(inc1 - cnt1)) * 32 + (as1 & 31);
        int room0 = a.K - (inc0 - cnt0), room1 = a.K - (inc1 - cnt1);   // entries beyond the capacity are dropped
#pragma unroll
        for (int ws = 0; ws < kTWords; ++ws) {
            uint32_t mm = m[ws];
            const Run &r = ws < 2 ? run0 : (ws < 4 ? run1 : run2);
            const Word x = word_of(r, ws & 1, g);
            while (mm) {                                                 // highest bit = first tile first
                const int b = 31 - __clz(mm);
                mm ^= 1u << b;
                const int32_t idx = (int32_t)I32[widx_of(x, b)];
                if (b & 1) {                                             // bits 3, 1 of a nibble: home particle 2t
                    if (room0 > 0) *e0p = idx;
                    e0p += 32;
                    --room0;
                } else {
                    if (room1 > 0) *e1p = idx;
                    e1p += 32;
                    --room1;
                }
            }
        }
        if (g == 0) {
            if (v0) a.cnt[as0] = tot0;
            if (v1) a.cnt[as1] = tot1;
        }
        wmax = max(wmax, (uint32_t)max(v0 ? tot0 : 0, v1 ? tot1 : 0));
    }
    return wmax;
}

// ------------------------------------------------------------------ neighbour kernel
// Blocks walk the groups with a grid stride.  On a dense grid the launcher gives every group its own block
// (the hardware balances them); on a sparse one (a sheet in a deep box leaves most groups empty) a resident set
// of blocks, so that an empty group costs two loads instead of a block launch.
template <bool RESIDENT>
__global__ void __launch_bounds__(kTThreads, kTBlocks)
tile_list_kernel(const __grid_constant__ sph_grid g, const __grid_constant__ TileArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *A = smem;
    uint32_t *I32 = reinterpret_cast<uint32_t *>(smem + kBytesA);
    uint32_t *Bv = reinterpret_cast<uint32_t *>(smem + kBytesA + kBytesI);
    Head *H = reinterpret_cast<Head *>(smem + kBytesA + kBytesI + kBytesBv);

    // positions far outside the box: single-shift semantics matter, the general path decides
    if (a.status->flags & (SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&a.status->flags, SPH_F_TILE_FALLBACK);
        return;
    }
    uint32_t wmax = 0;
    const uint32_t ngroups = g.ncode / 8u;
    for (uint32_t grp = blockIdx.x; grp < ngroups; grp += RESIDENT ? gridDim.x : ngroups) {
        const uint32_t c0 = grp * 8u;
        if (a.cell_start[c0 + 8] == a.cell_start[c0]) continue;         // no particle in the group
        if (RESIDENT) __syncthreads();                                   // the previous group's window is no longer read
        if (!tile_stage(g, c0, a, H, A, I32, Bv)) {
            if (threadIdx.x == 0) atomicOr(&a.status->flags, SPH_F_TILE_FALLBACK);
            break;
        }
        wmax = max(wmax, tile_cell(g, a, H, A, I32, Bv));
    }
    wmax = __reduce_max_sync(kFull, wmax);
    if ((threadIdx.x & 31) == 0 && wmax > 0) {
        if (wmax > *(volatile uint32_t *)&a.status->max_count) atomicMax(&a.status->max_count, wmax);
        if (wmax > (uint32_t)a.K) atomicOr(&a.status->flags, SPH_F_NBR_OVERFLOW);
    }
}

void tile_thresholds(const sph_grid *g, float *scale, float *thr_s, float *bw)
{
    // Everything in the scaled frame u = x / wmax (|u_c| <= 2 for a candidate, |u_p| <= 1 for a home particle; the
    // scaled threshold thr / wmax^2 is at most 1).  Error budget of d against the exact rsq / wmax^2 - thr':
    //   coordinates  fp32 staging: cell-relative conversion 1u w, shift conversion 2u w, their sum 2u w, the scaling
    //                2u: a separation component is off by err <= 16u, rsq by 2 sqrt(3) r err + 3 err^2 near the threshold;
    //   splitting    x = hi + lo + res, |res| <= 2^-21 |x| (half-precision rounding twice, subnormal lo included):
    //                the product x_c x_p misses lo lo + res terms, below 2.4e-6 per dimension, 1.5e-5 in -2 u_c . u_p;
    //   norms        |u|^2 in fp32 from the fp32 coordinates: 5u * 12; three halves carry it to 2^-30;
    //   accumulation the tensor core adds 16 exact products in fp32; allow 2^-22 of the largest partial sum
    //                (|u_c|^2 + |u_p|^2 + 2 |u_c . u_p| <= 12 + 3 + 12) per step -- truncation of an fp32 accumulator
    //                loses at most 2^-23 of it: 1.0e-4.
    // Use 4x the geometric part plus 1.5x the arithmetic part.
    double wmax = 0.0;
    for (int d = 0; d < 3; ++d) wmax = g->w[d] > wmax ? g->w[d] : wmax;
    const double u = 1.0 / 16777216.0;
    const float sf = (float)(1.0 / wmax);
    const double ss = (double)sf;                          // the scale the kernel multiplies by
    const double thr = g->thr * ss * ss, rl = sqrt(thr);
    const double err = 16.0 * u;
    const double geom = 2.0 * 1.7320508 * (rl + err) * err + 3.0 * err * err;
    const double arith = 1.5e-5 + 60.0 * u + 1.0e-4;
    const double band = 4.0 * geom + 1.5 * arith;
    float t = (float)(thr + band);
    t = nextafterf(t, INFINITY);
    *scale = sf;
    *thr_s = t;
    // hits with d >= -bw, i.e. rsq' >= thr' - bw, may lie outside: bw covers thr' - (thr - band) with margin
    const float w2 = (float)(((double)t - thr) + band);
    *bw = nextafterf(w2 * 1.0000002f, INFINITY);
}

TileArgs base_args(const sph_grid *g, const sph_buffers *b)
{
    TileArgs a = {};
    a.n = b->n;
    a.K = b->max_nbrs;
    a.cell_start = b->cell_start;
    a.rel4 = b->rel4;
    a.pos4 = b->pos4;
    a.nbr = b->nbr;
    a.cnt = b->cnt;
    a.status = b->status;
    a.perm = b->perm;
    a.n_owned = b->n_owned;
    tile_thresholds(g, &a.scale, &a.thr_s, &a.bw);
    return a;
}

}  // namespace

namespace sph_tiles_mma {

bool eligible(const sph_grid *g, const sph_buffers *b)
{
    if (!b->rel4 || !b->pos4 || !b->nbr || !b->cnt) return false;
    // four distinct layers per dimension: the window's outer layers and a tile's overshoot rows must be particles that
    // are two cells from the home cell under the minimum image too (with three layers they would be images of
    // neighbours, and the fp64 predicate of the in-band path could admit one twice)
    for (int d = 0; d < 3; ++d)
        if (g->ncl[d] < 4 || g->lb[d] < 1) return false;
    // the three lowest code bits are one bit of x, y, z: a group of 8 consecutive codes is 2 x 2 x 2 cells
    return (g->mask[0] & 7u) == 1u && (g->mask[1] & 7u) == 2u && (g->mask[2] & 7u) == 4u && (g->ncode % 8u) == 0u;
}

int launch_list(const sph_grid *g, const sph_buffers *b, cudaStream_t s)
{
    // per device: the attribute belongs to the function ON a device, and so does the SM count
    static bool configured[64] = {};
    static int sm_of[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev >= 0 && dev < 64 ? dev : 0;
    if (!configured[slot]) {
        cudaFuncSetAttribute(tile_list_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemList);
        cudaFuncSetAttribute(tile_list_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemList);
        cudaDeviceGetAttribute(&sm_of[slot], cudaDevAttrMultiProcessorCount, dev);
        if (sm_of[slot] <= 0) sm_of[slot] = 148;
        configured[slot] = true;
    }
    const TileArgs a = base_args(g, b);
    const unsigned groups = g->ncode / 8u, resident = (unsigned)(sm_of[slot] * kTBlocks);
    const bool sparse = (double)b->n < 24.0 * (double)groups;            // fewer than 3 particles per cell on average
    if (sparse && groups > resident) tile_list_kernel<true><<<resident, kTThreads, kSmemList, s>>>(*g, a);
    else tile_list_kernel<false><<<groups, kTThreads, kSmemList, s>>>(*g, a);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPH_OK : (int)e;
}

}  // namespace sph_tiles_mma
