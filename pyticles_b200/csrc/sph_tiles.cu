// pyticles_b200 -- cell-group ("tile") neighbour pass for sm_100a.
//
// One block works on a GROUP of 2 x 2 x 2 cells (eight consecutive block-Morton codes), one warp
// per cell.  The block stages the 4 x 4 x 4 cells around the group ONCE in shared memory as fp32
// positions in the group's frame (origin at the corner the eight cells share) with their squared norm, so a
// particle row is read from L2 8x instead of 27x, with coalesced loads.  The head of a block -- which 64 cells, how
// many particles each, where they go -- is ONE warp's work behind ONE barrier: a 64-byte row of the per-grid group
// table (sph_buffers.group_tab: the cell-code contributions of the 4 + 4 + 4 window layers; geometry only), three
// shuffles per cell code, the cell_start loads, a warp scan.
//
// A warp takes the particles of its cell up to 16 at a time.  Inside a pass, lane = q * P + p:
// particle p (of P) and candidate stream q of Q = 32 / P.  A lane walks the candidate pairs 2 q, 2 q + 1 of every 2 Q in the 9 window columns (3 cells
// each, contiguous in the staged window) around its cell and keeps its own list of hits -- no ballot, no popc, one predicated shared
// store per hit.  One test is the dot-product form
//     |c|^2 - 2 c.p  <  thr_out - |p|^2             3 FFMA + FSETP  (|c|^2 staged with the position, -2p and the
//                                                    right-hand side per lane)
// and, on a hit, the predicated store, pointer bump and running maximum for the error band: 7 instructions with the
// LDS.128 (r1: 12 with three subtractions, a multiply, two FFMA and a maximum of the accepted rsq).  The Q lists of a
// particle are then concatenated, in a fixed order, into its warp-transposed ELL row (the neighbour structure
// every other pass and the export use).
//
// Exactness is the one of the general kernel: rsq32 >= thr_out rejects, rsq32 < thr_out - bw accepts (bw: the
// rigorous fp32 error band, tile_thresholds); a lane one of whose hits fell between the two works the test value out
// again for each of its hits and re-decides those inside the band with the reference's fp64 predicate (pair_exact).  Cases outside the fixed capacities (> 64 particles in a cell, > 32 hits in one stream's
// list even at Q >= 4, > 1024 particles in the 64 cells of a window, positions far outside the box) raise
// SPH_F_TILE_FALLBACK and the general kernel redoes the pass.
//
// Round 2 tried three other formulations of this pass, each exact and tested, none faster on B200 (DESIGN.md section 5,
// profiles/r2_tile_experiments.txt): bit-mask bookkeeping instead of per-lane lists (fewer test instructions, but
// walking the masks keeps a third of the lanes busy: 5.4 ms), the pass body compiled per Q so that every stride is an
// immediate (11 % fewer instructions, but four copies of the unrolled column code miss the instruction cache: 4.9 ms),
// and a tensor-core pre-filter (sph_tiles_mma.cu, SPH_TILES=2: 5.7 ms).  What stayed is the dot-product form of the
// test with one copy of the pass: 4.26 ms against r1's 4.37.  The larger step came from the warp-state samples
// (profiles/summarise.py stalls): a third of them sat at barriers and on the long scoreboard, i.e. in serial latencies
// around the tests -- three barrier-separated head phases behind ~250 dependent integer instructions, one window row
// per loop trip, a status read before every warp's exit, every hit of a lane re-decided in fp64 when one candidate
// fell into the band.  Taking those out (see the head above, tile_stage, tile_pass, the kernel's last lines): 3.54 ms;
// 3.34 ms with the threshold as the comparison's operand and the leaner column set-up.
//
// Reference semantics (file:line into the reference tree):
//   pair predicate     neighbour_list.py:105-123,170-178
#include <stdlib.h>

#include "sph_device.cuh"
#include "sph_tiles.cuh"

namespace {

constexpr int kTWarps = 8;           // warps per block = cells per group
constexpr int kTThreads = kTWarps * 32;
constexpr int kTCap = 1024;          // staged candidates per group (64 cells).  1280 would be 44 KB per block: the
                                     // fifth block no longer fits an SM and the pass takes 4.61 instead of 4.26 ms
constexpr int kTPass = 16;           // particles of a cell per pass (Q = 32 / P >= 2 streams each); 8 when lists overflow
constexpr int kTPart = 64;           // particles per cell the tile path handles (cell width ~2 lattice planes: 8 .. 27)
constexpr int kTRow = 32;            // hits one lane (one stream of one particle) can hold
typedef uint16_t entry_t;            // a hit is the 16-bit shared address of the staged candidate
constexpr int kTStep = 32 * 2;       // bytes between consecutive hits of a lane: the lists of a warp are interleaved (hit k of lane l at
                                     // entry 32 k + l), so the lanes of a store fall into distinct banks (or share a word) at ANY mix of
                                     // list depths -- lists side by side at a 17-word stride cost 2.2 wavefronts per store
#ifndef SPH_TILE_BLOCKS
#define SPH_TILE_BLOCKS 5
#endif
constexpr int kTBlocks = SPH_TILE_BLOCKS;   // resident blocks per SM (38 KB of shared memory; 5: 48 registers)

constexpr uint32_t kFull = 0xffffffffu;

struct TileArgs {
    int n, K;
    const uint32_t *cell_start;
    const float *rel4;
    const double *pos4;
    int32_t *nbr;
    int32_t *cnt;
    sph_status *status;
    float thr_out, bw;       // hit: rsq32 < thr_out; hits with rsq32 >= thr_out - bw are settled in fp64
    int pass0;               // particles of a cell per pass to start with (16 or 8)
    int dot;                 // 1: dot-product form of the test, 0: difference form (wide cells)
    const int32_t *perm;
    int n_owned;             // > 0: cells of ghosts (original index >= n_owned) get empty rows
    const uint32_t *tab;     // sph_buffers.group_tab or nullptr: 16 words per group (group_table_kernel)
    float shift[12];         // (i - 2) * w[d] at [4 d + i]: fp32 frame shift of window layer i
};

// shared memory of a block: [S32 | I32 | B | Head]
struct Head {
    uint32_t off[68];        // exclusive scan of cnt (65 used; [65]: is there a particle in the group's own cells)
    uint32_t offb[68];       // the same as shared byte addresses of the staged candidates (S32 + 16 off)
    uint32_t start[64];      // first sorted particle of window cell (wz*4 + wy)*4 + wx
    uint32_t cnt[64];
    int gc[4];               // local cell coordinates of the group's base cell
    unsigned long long mbar; // SPH_TILE_TMA: transaction barrier of the bulk copies of the window
};

// SPH_TILE_TMA=1 (variant, not the default): the 64 cell segments of the window (8 particles x 16 bytes each,
// contiguous in the Morton-sorted rel4 array) are brought into shared memory by 64 one-dimensional bulk copies of the
// TMA unit (cp.async.bulk, SASS UBLKCP) completing on one mbarrier, and shifted into the group's frame IN shared memory
// afterwards, instead of LDG -> FADD -> STS per particle.  Measured on the 256^3 box: see profiles/r2_tile_experiments.txt.
#ifndef SPH_TILE_TMA
#define SPH_TILE_TMA 0
#endif

constexpr size_t kBytesS32 = sizeof(float4) * kTCap;
constexpr size_t kBytesI32 = sizeof(uint32_t) * kTCap;
constexpr size_t kBytesB = sizeof(entry_t) * kTWarps * 32 * kTRow;
constexpr size_t kSmemList = kBytesS32 + kBytesI32 + kBytesB + sizeof(Head);
static_assert(kTBlocks * (kSmemList + 1024) <= 227 * 1024, "blocks per SM");
static_assert(kBytesS32 + 1024 < 65536, "hits are kept as 16-bit shared addresses of the staged candidate");

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// Code contribution of cell layer c (before wrapping) of dimension d, or ~0u when the grid has no such layer: the cell
// code is additive over the dimensions, (block coordinate * block stride) << lbits plus the in-block Morton bits.
__device__ __forceinline__ uint32_t layer_part(const sph_grid &g, int d, int c)
{
    if (c < 0) {
        if (!g.wrap[d]) return ~0u;
        c += g.ncl[d];
    } else if (c >= g.ncl[d]) {
        if (!g.wrap[d]) return ~0u;
        c -= g.ncl[d];
    }
    const uint32_t stride = d == 0 ? 1u : (d == 1 ? g.nblk[0] : g.nblk[0] * g.nblk[1]);
    return ((((uint32_t)c >> g.lb[d]) * stride) << g.lbits) | pdep32((uint32_t)c & ((1u << g.lb[d]) - 1u), g.mask[d]);
}

// What a block needs to know about its group and depends on the grid only: word 4 d + i of a group's 16 is the code
// contribution of window layer i of dimension d, words 12..14 the cell coordinates of its base cell.
__global__ void __launch_bounds__(256)
group_table_kernel(const __grid_constant__ sph_grid g, uint32_t *__restrict__ tab)
{
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x, grp = gt >> 4, l = gt & 15u;
    if (grp >= g.ncode / 8u) return;
    int cc[3];
    cell_coords(g, grp * 8u, cc[0], cc[1], cc[2]);
    uint32_t out = 0u;
    if (l < 12u) out = layer_part(g, (int)(l >> 2), cc[l >> 2] + (int)(l & 3u) - 1);
    else if (l < 15u) out = (uint32_t)cc[l - 12u];
    tab[gt] = out;
}

// ------------------------------------------------------------------ window of a group
// Fills Head (window cells, their scan, frame shifts, base coordinates) and stages the window:
// S32 = fp32 position in the group's frame + its squared norm, I32 = sorted index.  Returns the number of
// staged candidates (> kTCap: nothing was staged).
__device__ __forceinline__ uint32_t tile_stage(const sph_grid &g, uint32_t c0, const TileArgs &a, Head *H,
                                               float4 *S32, uint32_t *I32, uint32_t phase)
{
    const int t = threadIdx.x;
    if (t < 32) {
        // ONE warp fills the head, so that a single barrier stands between the start of the block and the staging of the
        // window (the table, the look-ups and the scan used to be three phases with a barrier after each: the other
        // warps of the block spent a fifth of its lifetime waiting at them).
        // The cell code is additive over the dimensions: (block coordinate * block stride) << lbits plus the
        // in-block Morton bits.  Lane 4 d + i works out the contribution of window layer i of dimension d;
        // the 64 cell codes are then three shuffles and two adds each, two cells (2 t, 2 t + 1) per lane.
        uint32_t part = ~0u;
        if (a.tab) {
            // (one 64-byte row of the per-grid table instead of ~150 dependent integer instructions)
            const uint32_t v = t < 16 ? a.tab[2u * c0 + (uint32_t)t] : 0u;
            if (t < 12) part = v;
            else if (t < 15) H->gc[t - 12] = (int)v;
        } else {
            const int d = (t >> 2) % 3, i = t & 3;
            uint32_t bx, by, bz;
            block_coords(g, c0 >> g.lbits, bx, by, bz);
            const int ccd = (int)(((d == 0 ? bx : (d == 1 ? by : bz)) << g.lb[d]) | pext32(c0, g.mask[d]));
            if (t < 12) {
                part = layer_part(g, d, ccd + i - 1);
                if (i == 0) H->gc[d] = ccd;
            }
        }
        uint32_t st[2], cn[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = 2 * t + j;
            const uint32_t px = __shfl_sync(kFull, part, c & 3), py = __shfl_sync(kFull, part, 4 + ((c >> 2) & 3)),
                           pz = __shfl_sync(kFull, part, 8 + (c >> 4));
            st[j] = 0; cn[j] = 0;
            if (px != ~0u && py != ~0u && pz != ~0u) {
                const uint32_t code = px + py + pz;
                st[j] = a.cell_start[code];
                cn[j] = a.cell_start[code + 1];
            }
        }
        const uint32_t v0 = cn[0] - st[0], v1 = cn[1] - st[1];
        uint32_t inc = v0 + v1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(kFull, inc, o);
            if (t >= o) inc += x;
        }
        const uint32_t ex = inc - (v0 + v1);
        // the group's own cells are window cells 21, 22, 25, 26, 37, 38, 41, 42: is there a particle in any of them?
        const uint32_t own = ((0x00141400u >> t) & 1u) ? v1 : (((0x00282800u >> t) & 1u) ? v0 : 0u);
        const bool any = __any_sync(kFull, own != 0u);
        *reinterpret_cast<uint2 *>(&H->start[2 * t]) = make_uint2(st[0], st[1]);
        *reinterpret_cast<uint2 *>(&H->cnt[2 * t]) = make_uint2(v0, v1);
        *reinterpret_cast<uint2 *>(&H->off[2 * t]) = make_uint2(ex, ex + v0);
        const uint32_t s32a = smem_u32(S32);
        *reinterpret_cast<uint2 *>(&H->offb[2 * t]) = make_uint2(s32a + 16u * ex, s32a + 16u * (ex + v0));
        if (t == 31) { H->off[64] = inc; H->off[65] = any ? 1u : 0u; H->offb[64] = s32a + 16u * inc; }
    }
    __syncthreads();
    const uint32_t total = H->off[64];
    if (H->off[65] == 0u) return 0u;                                     // nothing to do for this group
    if (total > (uint32_t)kTCap) return total;
    // warp w stages window cells 8w .. 8w+7, four cells per pass (8 lanes each)
    const int w = t >> 5, lane = t & 31;
    const float4 *rel = reinterpret_cast<const float4 *>(a.rel4);
#if SPH_TILE_TMA
    {
        const uint32_t mbar = smem_u32(&H->mbar);
        if (t == 0) {
            // the window was last touched through the generic proxy (the previous group's tests): order it before
            // the asynchronous writes, then arm the barrier with the bytes to expect
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(total * 16u) : "memory");
        }
        __syncthreads();
        if (t < 64 && H->cnt[t]) {
            const uint32_t dst = smem_u32(S32 + H->off[t]);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(rel + H->start[t]), "r"(H->cnt[t] * 16u), "r"(mbar) : "memory");
        }
        // every thread waits for the phase to complete (parity alternates from group to group)
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
    }
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int wc = w * 8 + pass * 4 + (lane >> 3);
        const uint32_t st = H->start[wc], cn = H->cnt[wc], dst = H->off[wc];
        const float fx = a.shift[wc & 3], fy = a.shift[4 + ((wc >> 2) & 3)], fz = a.shift[8 + (wc >> 4)];
        for (uint32_t k = lane & 7; k < cn; k += 8) {
            const float4 p = S32[dst + k];
            const float x = p.x + fx, y = p.y + fy, z = p.z + fz;
            S32[dst + k] = make_float4(x, y, z, fmaf(z, z, fmaf(y, y, x * x)));
            I32[dst + k] = st + k;
        }
    }
#else
    (void)phase;
    {
        // lane (c, k) = (lane >> 3, lane & 7) takes particles k, k + 8, ... of window cells 8w + c and 8w + 4 + c.  The
        // first two rows of both cells are requested before any is used (four loads in flight instead of one per loop trip)
        const uint32_t k0 = lane & 7;
        uint32_t st[2], cn[2], dst[2];
        float fx[2], fy[2], fz[2];
        float4 p[4];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int wc = w * 8 + pass * 4 + (lane >> 3);
            st[pass] = H->start[wc]; cn[pass] = H->cnt[wc]; dst[pass] = H->off[wc];
            fx[pass] = a.shift[wc & 3]; fy[pass] = a.shift[4 + ((wc >> 2) & 3)]; fz[pass] = a.shift[8 + (wc >> 4)];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t k = k0 + 8u * (uint32_t)(i & 1);
            if (k < cn[i >> 1]) p[i] = __ldg(rel + st[i >> 1] + k);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = i >> 1;
            const uint32_t k = k0 + 8u * (uint32_t)(i & 1);
            if (k < cn[c]) {
                const float x = p[i].x + fx[c], y = p[i].y + fy[c], z = p[i].z + fz[c];
                S32[dst[c] + k] = make_float4(x, y, z, fmaf(z, z, fmaf(y, y, x * x)));
                I32[dst[c] + k] = st[c] + k;
            }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
            for (uint32_t k = k0 + 16u; k < cn[c]; k += 8) {             // cells of more than 16 particles
                const float4 q = __ldg(rel + st[c] + k);
                const float x = q.x + fx[c], y = q.y + fy[c], z = q.z + fz[c];
                S32[dst[c] + k] = make_float4(x, y, z, fmaf(z, z, fmaf(y, y, x * x)));
                I32[dst[c] + k] = st[c] + k;
            }
    }
#endif
    __syncthreads();
    return total;
}

// One warp = one home cell of the group, worked on kTPass particles at a time.
struct HomeCell {
    int P;                   // particles in the cell
    uint32_t cs;             // first sorted particle of the cell
    uint32_t c0;             // window index of its first particle
    int hx, hy, hz;
};

__device__ __forceinline__ HomeCell home_cell(const sph_grid &g, const Head *H, int w)
{
    HomeCell h;
    h.hx = w & 1; h.hy = (w >> 1) & 1; h.hz = w >> 2;
    const int wc = ((1 + h.hz) * 4 + (1 + h.hy)) * 4 + (1 + h.hx);
    // the last group of an odd layer count is half empty: its window cell is the periodic image of layer 0
    const bool exists = H->gc[0] + h.hx < g.ncl[0] && H->gc[1] + h.hy < g.ncl[1] && H->gc[2] + h.hz < g.ncl[2];
    h.P = exists ? (int)H->cnt[wc] : 0;
    h.cs = H->start[wc];
    h.c0 = H->off[wc];
    return h;
}

#define SPH_LDS4(X, Y, Z, W, ADDR) \
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(X), "=f"(Y), "=f"(Z), "=f"(W) : "r"(ADDR))

// Sorted index of the staged candidate whose shared byte address (16 bits: the window lies below 64 KB) a hit list
// holds: candidate c = (ptr - S32) / 16 has its index at I32 + 4 c = ptr / 4 + (I32 - S32 / 4) -- one shift-and-add
// (LEA.HI) in front of the load, with i32c = I32 - S32 / 4 in a register.
__device__ __forceinline__ uint32_t hit_index(uint32_t raw, uint32_t i32c)
{
    uint32_t j;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(j) : "r"((raw >> 2) + i32c));
    return j;
}

// The quantity a candidate is judged by, against the home particle's threshold T:
// DOT    d = |c|^2 - 2 c.p against T = thr_out - |p|^2, (px2, py2, pz2) = -2 p: three FFMA on the staged (c, |c|^2).  Its
//        rounding error grows with the SQUARE of the coordinates, so grids with a cell much wider than the list radius in
//        some dimension (a sheet in a deep box: coarse z cells) take
// !DOT   d = |c - p|^2 against T = thr_out, (px2, py2, pz2) = p: three subtractions, a multiply and two FFMA.
template <bool DOT>
__device__ __forceinline__ float test_value(float px2, float py2, float pz2, float cx, float cy, float cz, float cn)
{
    if (DOT) return fmaf(cz, pz2, fmaf(cy, py2, fmaf(cx, px2, cn)));
    const float dx = cx - px2, dy = cy - py2, dz = cz - pz2;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// One candidate against one home particle.  A hit (d < T, i.e. rsq32 < thr_out) is kept as the 16-bit shared address of
// the staged candidate; `hmax` follows the largest d among the hits: only a hit with d >= T - bw can lie outside, and
// the caller settles those in fp64 (a miss needs nothing, however close).  Nothing else happens per test: the
// comparison takes the threshold as its second operand, so neither it nor the band costs an instruction of its own.
// SELF: the column holds the particle itself.
template <bool DOT, bool CHECK, bool SELF>
__device__ __forceinline__ void test_one(uint32_t ptr, uint32_t selfptr, float px2, float py2, float pz2, float T,
                                         float cx, float cy, float cz, float cn, uint32_t &lp, uint32_t lp_lim,
                                         float &hmax, bool &over)
{
    const float d = test_value<DOT>(px2, py2, pz2, cx, cy, cz, cn);
    if (d < T && (!SELF || ptr != selfptr)) {
        hmax = fmaxf(hmax, d);
        if (CHECK && lp >= lp_lim) {
            over = true;
        } else {
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(lp), "h"((unsigned short)ptr) : "memory");
            lp += (uint32_t)kTStep;
        }
    }
}

// One candidate stream of one window column: stream q of Q takes the candidate PAIRS 2 q, 2 q + 1 (+ 2 Q, + 4 Q, ...) from
// the column's start, i.e. ptr, ptr + 16, then ptr + step2 (= 32 Q bytes), ... below pend -- two adjacent candidates per
// trip.  (Pairs rather than every Q-th candidate: the two rows a stream takes from a cell of eight then share a
// 128-byte line of the fp64 arrays, and the rows of a particle list its neighbours stream by stream -- the lanes of a
// cell, which sweep their rows together in the density and force passes, ask L1 for half as many lines.)
template <bool DOT, bool CHECK, bool SELF>
__device__ __forceinline__ void test_column(uint32_t ptr, uint32_t pend, uint32_t step, uint32_t selfptr, float px2,
                                            float py2, float pz2, float T, uint32_t &lp, uint32_t lp_lim, float &hmax,
                                            bool &over)
{
    float ax, ay, az, an, bx, by, bz, bn;
    for (; ptr + 16u < pend; ptr += step) {                              // two loads in flight
        SPH_LDS4(ax, ay, az, an, ptr);
        SPH_LDS4(bx, by, bz, bn, ptr + 16u);
        test_one<DOT, CHECK, SELF>(ptr, selfptr, px2, py2, pz2, T, ax, ay, az, an, lp, lp_lim, hmax, over);
        test_one<DOT, CHECK, SELF>(ptr + 16u, selfptr, px2, py2, pz2, T, bx, by, bz, bn, lp, lp_lim, hmax, over);
    }
    if (ptr < pend) {
        SPH_LDS4(ax, ay, az, an, ptr);
        test_one<DOT, CHECK, SELF>(ptr, selfptr, px2, py2, pz2, T, ax, ay, az, an, lp, lp_lim, hmax, over);
    }
}

// One pass: P <= 32 / Q particles of the home cell (first at window index c0, sorted index cs), lane = q * P + p
// (particle p, candidate stream q of Q).  Returns the longest row written, or ~0u when a stream's list overflowed.
// (Compiling the pass per Q makes every stride an immediate and saves 11 % of the instructions, but the four
// copies of the unrolled column code no longer fit the instruction cache: 4.89 against 4.37 ms, profiles/r2_tile_*.)
template <bool DOT>
__device__ __forceinline__ uint32_t tile_pass(const sph_grid &g, const TileArgs &a, const float4 *S32,
                                              const uint32_t *I32, entry_t *B, const uint32_t *offh, uint32_t c0,
                                              uint32_t cs, int P, int Q, int lane)
{
    const uint32_t step = (uint32_t)Q * 32u;                             // a trip: the next pair of this stream
    const int q = (int)(((float)lane + 0.5f) * __frcp_rn((float)P));     // lane / P: never within 1/64 of an integer
    const int p = lane - q * P;
    const bool active = q < Q;
    const uint32_t selfc = c0 + (uint32_t)p, asorted = cs + (uint32_t)p;
    const uint32_t s32a = smem_u32(S32), selfptr = s32a + selfc * 16u, i32c = smem_u32(I32) - (s32a >> 2);
    const float4 hp = S32[selfc];
    const float px2 = DOT ? -2.0f * hp.x : hp.x, py2 = DOT ? -2.0f * hp.y : hp.y, pz2 = DOT ? -2.0f * hp.z : hp.z,
                T = DOT ? a.thr_out - hp.w : a.thr_out, Tsure = T - a.bw;
    const uint32_t lp0 = smem_u32(B), lp_lim = lp0 + (uint32_t)kTStep * kTRow, q16 = (uint32_t)q * 32u;
    uint32_t lp = lp0;
    float hmax = -INFINITY;
    bool over = false;
    if (active) {
#pragma unroll
        for (int col = 0; col < 9; ++col) {                              // unrolled: 4.43 ms against 4.55 ms as a loop
            // (byte addresses of the column's first candidate and of its end)
            const uint32_t sb = offh[((col / 3) * 4 + col % 3) * 4], pend = offh[((col / 3) * 4 + col % 3) * 4 + 3];
            const uint32_t pbeg = sb + q16;
            // this lane tests at most n / Q + 2 <= n / 2 + 2 of the column's n candidates (Q >= 2): with room for that
            // many hits (kTStep bytes each: 2 * 16 n + 2 kTStep) the loop needs no capacity test
            const bool room = lp + ((pend - sb) << 1) + 2u * (uint32_t)kTStep <= lp_lim;
            if (col == 4) {                                              // the column that holds the particle itself
                if (room) test_column<DOT, false, true>(pbeg, pend, step, selfptr, px2, py2, pz2, T, lp, lp_lim, hmax, over);
                else test_column<DOT, true, true>(pbeg, pend, step, selfptr, px2, py2, pz2, T, lp, lp_lim, hmax, over);
            } else {
                if (room) test_column<DOT, false, false>(pbeg, pend, step, selfptr, px2, py2, pz2, T, lp, lp_lim, hmax, over);
                else test_column<DOT, true, false>(pbeg, pend, step, selfptr, px2, py2, pz2, T, lp, lp_lim, hmax, over);
            }
        }
    }
    if (__any_sync(kFull, over)) return ~0u;
    // rare: a hit of this lane lies in the fp32 error band -> d is worked out again for every hit of the lane (the same
    // operations on the same operands give the same bits) and the hits inside the band are decided by the reference's
    // fp64 predicate on the fp64 rows
    if (hmax >= Tsure) {
        const int nl = (int)((lp - lp0) / (uint32_t)kTStep);
        int m = 0;
        for (int k = 0; k < nl; ++k) {
            const entry_t raw = B[k * 32];
            float cx, cy, cz, cn;
            SPH_LDS4(cx, cy, cz, cn, (uint32_t)raw);
            const float d = test_value<DOT>(px2, py2, pz2, cx, cy, cz, cn);
            if (d < Tsure || pair_exact(g, a.pos4, (int)asorted, (int)hit_index(raw, i32c))) B[32 * m++] = raw;
        }
        lp = lp0 + (uint32_t)kTStep * (uint32_t)m;
    }
    const int cntl = (int)((lp - lp0) / (uint32_t)kTStep);
    // concatenate the Q lists of a particle: offsets by a fixed-order walk over the streams
    int offq = 0, tot = 0;
#pragma unroll 1
    for (int k = 0; k < Q; ++k) {                                        // Q is uniform across the warp
        const int v = __shfl_sync(kFull, cntl, (k * P + p) & 31);
        offq += k < q ? v : 0;
        tot += v;
    }
    if (!active) return 0u;
    int32_t *erow = a.nbr + ((size_t)(asorted >> 5) * (size_t)a.K + (size_t)offq) * 32 + (asorted & 31);
    const int nw = min(cntl, a.K - offq);                                // entries beyond the capacity are dropped
    int k = 0;
    for (; k + 4 <= nw; k += 4) {                                        // four independent look-up chains in flight
        const uint32_t b0 = B[k * 32], b1 = B[k * 32 + 32], b2 = B[k * 32 + 64], b3 = B[k * 32 + 96];
        const uint32_t j0 = hit_index(b0, i32c), j1 = hit_index(b1, i32c), j2 = hit_index(b2, i32c),
                       j3 = hit_index(b3, i32c);
        erow[k * 32] = (int32_t)j0; erow[(k + 1) * 32] = (int32_t)j1;
        erow[(k + 2) * 32] = (int32_t)j2; erow[(k + 3) * 32] = (int32_t)j3;
    }
#pragma unroll 1
    for (; k < nw; ++k) erow[k * 32] = (int32_t)hit_index(B[k * 32], i32c);
    if (q == 0) a.cnt[asorted] = tot;
    return (uint32_t)tot;
}

// ------------------------------------------------------------------ one cell of a staged group (one warp)
// Returns the longest row it wrote (0 when it gave up and raised SPH_F_TILE_FALLBACK).
template <bool DOT>
__device__ __forceinline__ uint32_t tile_cell(const sph_grid &g, const TileArgs &a, const Head *H, const float4 *S32,
                                              const uint32_t *I32, entry_t *Bblock)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const HomeCell hc = home_cell(g, H, w);
    if (hc.P == 0) return 0u;
    if (hc.P > kTPart) {
        if (lane == 0) atomicOr(&a.status->flags, SPH_F_TILE_FALLBACK);
        return 0u;
    }
    if (a.n_owned > 0 && !g.wrap[0]) {
        // slab decomposition: the first and the last local x layer hold the ghosts (sph_grid_restrict_x keeps one
        // ghost layer on each side of the owned ones); no rows are built for them
        const int cx = H->gc[0] + hc.hx;
        if (cx == 0 || cx == g.ncl[0] - 1) {
            for (int k = lane; k < hc.P; k += 32) a.cnt[hc.cs + k] = 0;
            return 0u;
        }
    }
    entry_t *B = Bblock + w * 32 * kTRow + lane;                         // this lane's list of hits: B[32 k]
    const uint32_t *offh = H->offb + (hc.hz * 4 + hc.hy) * 4 + hc.hx;
    uint32_t wmax = 0;

    // Streams per particle: Q = 32 / P (8 at most).  Long rows (the default
    // Verlet tolerance gives ~47 neighbours) overflow a stream's list at Q = 2: the launcher then starts at 8
    // particles per pass (Q >= 4); a cell that overflows at 16 is redone at 8, then at 4, before the general kernel
    // is asked.
    // A pass of P particles costs every lane 1 / Q = 1 / floor(32 / P) of the window: 1/4 for P = 7..8, 1/3 for 9..10,
    // 1/2 for 11..16.  So at most 10 particles go into one pass and a longer remainder is taken 8 at a time: a cell of 12
    // costs 1/4 + 1/8 instead of 1/2 -- and the cells above 10 are the ones their block waits for.
    int pass = a.pass0;
    int p0 = 0;
#pragma unroll 1
    while (p0 < hc.P) {
        const int rem = hc.P - p0;
        const int P = pass > 8 ? (rem <= 10 ? rem : 8) : min(pass, rem);
        const uint32_t c0 = hc.c0 + (uint32_t)p0, cs = hc.cs + (uint32_t)p0;
        const uint32_t r = tile_pass<DOT>(g, a, S32, I32, B, offh, c0, cs, P, P <= 4 ? 8 : 32 / P, lane);
        if (r == ~0u) {
            if (pass > 4) {
                pass = pass > 8 ? 8 : 4;
                p0 = 0;                                                  // start the cell again
                continue;
            }
            if (lane == 0) atomicOr(&a.status->flags, SPH_F_TILE_FALLBACK);
            return 0u;
        }
        wmax = max(wmax, r);
        p0 += P;
        __syncwarp();                                                    // lists are reused by the next pass
    }
    return wmax;
}

// ------------------------------------------------------------------ neighbour kernel
// Blocks walk the groups with a grid stride.  On a dense grid the launcher gives every group its own block
// (the hardware balances them: 4.25 ms against 4.9 ms for resident blocks on 256^3); on a sparse one (a
// sheet in a deep box leaves most groups empty) a resident set of blocks, so that an empty group costs
// two loads instead of a block launch (0.71 -> 0.60 ms on the 1024^2 sheet).
template <bool RESIDENT, bool DOT>
__global__ void __launch_bounds__(kTThreads, kTBlocks)
tile_list_kernel(const __grid_constant__ sph_grid g, const __grid_constant__ TileArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *S32 = reinterpret_cast<float4 *>(smem);
    uint32_t *I32 = reinterpret_cast<uint32_t *>(smem + kBytesS32);
    entry_t *B = reinterpret_cast<entry_t *>(smem + kBytesS32 + kBytesI32);
    Head *H = reinterpret_cast<Head *>(smem + kBytesS32 + kBytesI32 + kBytesB);

    // Positions far outside the box: single-shift semantics matter, the general path decides.  A block with one group
    // of its own does not wait for the flags before it starts (one dependent round trip to L2 less per block: what it
    // writes is finite garbage in that case and the general kernel behind overwrites every row); block 0 raises the
    // fallback at the end.
    if (smem_u32(S32) + kBytesS32 > 65536u ||
        (RESIDENT && (a.status->flags & (SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE)))) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&a.status->flags, SPH_F_TILE_FALLBACK);
        return;
    }
    uint32_t wmax = 0, phase = 0;
#if SPH_TILE_TMA
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&H->mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#endif
    const uint32_t ngroups = g.ncode / 8u;
    for (uint32_t grp = blockIdx.x; grp < ngroups; grp += RESIDENT ? gridDim.x : ngroups) {
        const uint32_t c0 = grp * 8u;
        // no particle in the group: a resident block looks before it stages (most groups of a sparse grid are
        // empty); a block with one group learns it from the window counts, without a round trip of its own
        if (RESIDENT && a.cell_start[c0 + 8] == a.cell_start[c0]) continue;
        if (RESIDENT) __syncthreads();                                   // the previous group's window is no longer read
        const uint32_t total = tile_stage(g, c0, a, H, S32, I32, phase);
        if (total == 0u) continue;
        if (total <= (uint32_t)kTCap) phase ^= 1u;                       // (SPH_TILE_TMA: one barrier phase per staged window)
        if (total > (uint32_t)kTCap) {
            if (threadIdx.x == 0) atomicOr(&a.status->flags, SPH_F_TILE_FALLBACK);
            break;
        }
        wmax = max(wmax, tile_cell<DOT>(g, a, H, S32, I32, B));
    }
    wmax = __reduce_max_sync(kFull, wmax);
    // (only rows longer than the capacity report their length: a look at status->max_count before every warp's exit
    // is a round trip to L2 during which the block keeps its shared memory)
    if ((threadIdx.x & 31) == 0 && wmax > (uint32_t)a.K &&
        !(*(volatile uint32_t *)&a.status->flags & (SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE))) {   // (those rows do not count)
        atomicMax(&a.status->max_count, wmax);
        atomicOr(&a.status->flags, SPH_F_NBR_OVERFLOW);
    }
    if (!RESIDENT && blockIdx.x == 0 && threadIdx.x == 0 &&
        (*(volatile uint32_t *)&a.status->flags & (SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE)))
        atomicOr(&a.status->flags, SPH_F_TILE_FALLBACK);
}

// Thresholds of d = rsq32 - thr_out for the two forms of the test; returns whether the dot-product form is used.
bool tile_thresholds(const sph_grid *g, float *tout, float *bw)
{
    // Coordinates are fp32 in the GROUP's frame, |coordinate| <= 2 w: a staged coordinate carries at most 5u w
    // (cell-relative conversion 1u w, shift conversion 2u w, their sum 2u w), so a separation component is off by
    // err <= 10u w (11u w with the subtraction of the difference form) and rsq by 2 sqrt(3) r err + 3 err^2 near the
    // threshold.  On top of that the arithmetic of d:
    //   difference form  d = dx^2 - thr_out + dy^2 + dz^2: three FFMA on magnitudes <= thr near the threshold: 8u thr;
    //   dot-product form d = |c|^2 + (|p|^2 - thr_out) - 2 c.p: |c|^2 <= 12 w^2 (3 operations), |p|^2 <= 3 w^2 (3),
    //                    their sum with -thr_out (2), three FFMA on magnitudes <= 27 w^2 + thr: below
    //                    u (36 + 9 + 2 * 15 + 3 * 27) w^2 + 5u thr = 156u w^2 + 5u thr -- it grows with the SQUARE of the
    //                    cell width, so it is only used while that stays below a thousandth of the threshold (cells
    //                    up to five list radii wide; a sheet in a deep box has z cells a hundred radii wide).
    // Use 4x the sum, as sph_grid_plan does for the per-cell frame.
    double wmax = 0.0;
    for (int d = 0; d < 3; ++d) wmax = g->w[d] > wmax ? g->w[d] : wmax;
    const double u = 1.0 / 16777216.0;
    const double rl = sqrt(g->thr);
    const double arith_dot = 156.0 * u * wmax * wmax + 5.0 * u * g->thr;
    const bool dot = 4.0 * arith_dot < 1.0e-3 * g->thr;
    const double err = (dot ? 10.0 : 11.0) * u * wmax;
    const double band = 4.0 * (2.0 * 1.7320508 * (rl + err) * err + 3.0 * err * err + (dot ? arith_dot : 8.0 * u * g->thr));
    float b = (float)(g->thr + band);
    b = nextafterf(b, INFINITY);
    *tout = b;
    // hits with d >= -bw, i.e. rsq32 >= thr_out - bw, may lie outside: bw covers thr_out - (thr - band) with margin
    float w2 = (float)(((double)b - g->thr) + band);
    *bw = nextafterf(w2 * 1.0000002f, INFINITY);
    return dot;
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
    a.tab = b->group_tab;
    for (int d = 0; d < 3; ++d)
        for (int i = 0; i < 4; ++i) a.shift[4 * d + i] = (float)((double)(i - 2) * g->w[d]);
    a.dot = tile_thresholds(g, &a.thr_out, &a.bw) ? 1 : 0;
    // expected neighbours per particle at the mean density of the local grid: a stream of Q = 2 holds 32 hits.
    // (The owned particles over the owned layers when there are ghosts: the count must not depend on the capacity
    // of the ghost region, or the row order -- and with it the last bits of every sum -- would.)
    const bool slab = b->n_owned > 0 && !g->wrap[0] && g->ncl[0] > 2;
    const double vol = ((slab ? g->ncl[0] - 2 : g->ncl[0]) * g->w[0]) * (g->ncl[1] * g->w[1]) * (g->ncl[2] * g->w[2]);
    const double np = slab ? (double)b->n_owned : (double)b->n;
    const double expect = vol > 0.0 ? 4.18879 * g->thr * sqrt(g->thr) * np / vol : 0.0;
    a.pass0 = expect > 38.0 ? 8 : kTPass;
    if (const char *e = getenv("SPH_TILE_PASS0")) a.pass0 = atoi(e) > 0 ? atoi(e) : a.pass0;   // (experiments)
    return a;
}

}  // namespace

namespace sph_tiles {

bool grid_has_groups(const sph_grid *g)
{
    for (int d = 0; d < 3; ++d)
        if (g->ncl[d] < 3 || g->lb[d] < 1) return false;
    // the three lowest code bits are one bit of x, y, z: a group of 8 consecutive codes is 2 x 2 x 2 cells
    return (g->mask[0] & 7u) == 1u && (g->mask[1] & 7u) == 2u && (g->mask[2] & 7u) == 4u && (g->ncode % 8u) == 0u;
}

bool eligible(const sph_grid *g, const sph_buffers *b)
{
    const char *e = getenv("SPH_TILES");                 // SPH_TILES=0: general kernel only (tests, A/B timing)
    if ((e && atoi(e) == 0) || !b->rel4 || !b->pos4 || !b->nbr || !b->cnt) return false;
    return grid_has_groups(g);
}

int64_t group_tab_elems(const sph_grid *g)
{
    return grid_has_groups(g) ? (int64_t)(g->ncode / 8u) * 16 : 0;
}

int fill_group_table(const sph_grid *g, uint32_t *tab, cudaStream_t s)
{
    const int64_t lanes = group_tab_elems(g);
    if (lanes == 0) return SPH_OK;
    group_table_kernel<<<(unsigned)((lanes + 255) / 256), 256, 0, s>>>(*g, tab);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPH_OK : (int)e;
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
        cudaFuncSetAttribute(tile_list_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemList);
        cudaFuncSetAttribute(tile_list_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemList);
        cudaFuncSetAttribute(tile_list_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemList);
        cudaFuncSetAttribute(tile_list_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemList);
        cudaDeviceGetAttribute(&sm_of[slot], cudaDevAttrMultiProcessorCount, dev);
        if (sm_of[slot] <= 0) sm_of[slot] = 148;
        configured[slot] = true;
    }
    const TileArgs a = base_args(g, b);
    const unsigned groups = g->ncode / 8u, resident = (unsigned)(sm_of[slot] * kTBlocks);
    const bool sparse = (double)b->n < 24.0 * (double)groups;            // fewer than 3 particles per cell on average
    // (one instantiation per form: only the one launched is ever fetched, so the instruction cache sees one copy)
    if (sparse && groups > resident) {
        if (a.dot) tile_list_kernel<true, true><<<resident, kTThreads, kSmemList, s>>>(*g, a);
        else tile_list_kernel<true, false><<<resident, kTThreads, kSmemList, s>>>(*g, a);
    } else {
        if (a.dot) tile_list_kernel<false, true><<<groups, kTThreads, kSmemList, s>>>(*g, a);
        else tile_list_kernel<false, false><<<groups, kTThreads, kSmemList, s>>>(*g, a);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPH_OK : (int)e;
}

}  // namespace sph_tiles
