// pyticles_b200 -- hand-written sm_100a kernels of the SPH step hot path + their C ABI.
//
// Design in one paragraph (DESIGN.md has the long form).  Particles are binned into a
// periodic cell grid (cell width >= list radius), cells are numbered along a generalised
// Morton curve, and a counting sort gives a Morton-ordered permutation.  The particle
// state is gathered once per evaluation into packed, 32-byte rows in that order
// (pos4 = x y z m, vel4 = vx vy vz press/rho^2, rel4 = fp32 cell-relative position).  The
// neighbour pass runs one warp per cell: candidates from the 27 surrounding cells are
// staged in shared memory, every lane tests one candidate against the cell's particles in
// fp32 with a rigorous error band, the rare in-band candidates are decided by the
// reference's own fp64 predicate, and accepted neighbours are compacted with warp ballots
// into warp-transposed ELL rows.  Density/EOS and force passes run one thread per
// particle over those rows with fp64 accumulation in registers: no atomics, fixed
// summation order, bit-reproducible results.
//
// Reference semantics restated here (file:line into the reference tree):
//   minimum image      neighbour_list.py:105-123
//   pair predicate     neighbour_list.py:170-178   rsq < cutoff^2 + tolerance^2, strict
//   Lucy kernel        spkernel.py:86-118
//   density, EOS       properties.py:38-49,63-120
//   pressure force     forces.py:38-42,327-368 (2-D :246-274; cohesive :371-405)
//   list maintenance   neighbour_list.py:191-234
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "pyticles_b200.h"

#include "sph_device.cuh"
#include "sph_tiles.cuh"

// Compile-time tunables of the density / force passes.  tools/variant_sweep.py builds one library per setting
// and times them on the bench workload; the defaults are the fastest of those sweeps on B200 (256^3 box,
// profiles/r1f_variant_sweep_gen*.txt): 128-thread blocks capped at 72 registers (7 blocks = 28 warps per SM
// instead of 24 / 16), the force pass gathering one neighbour per stage instead of two, indices fetched two
// stages ahead.  Every setting evaluates a particle's row in the same order: the results are bit-identical.
//   SPH_PP_BLOCK     threads per block            SPH_DENS_MINB / SPH_FORCE_MINB   __launch_bounds__ min blocks
//   SPH_ROW_U / SPH_ROW_UF   neighbours gathered per pipeline stage (density / force)
//   SPH_IDX_AHEAD / SPH_IDX_AHEAD_F   pipeline stages the neighbour indices are fetched ahead of their rows
// Other variants of these passes were built, measured and taken out again (cache hints, double-buffered rows,
// per-SM chunk queues, lane-pair gathers of interleaved rows, in-warp neighbours by shuffle): DESIGN.md section 5.
#ifndef SPH_PP_BLOCK
#define SPH_PP_BLOCK 128
#endif
#ifndef SPH_DENS_MINB
#define SPH_DENS_MINB 7
#endif
#ifndef SPH_FORCE_MINB
#define SPH_FORCE_MINB 7
#endif
#ifndef SPH_ROW_U
#define SPH_ROW_U 4
#endif
#ifndef SPH_ROW_UF
#define SPH_ROW_UF 1
#endif
#ifndef SPH_IDX_AHEAD
#define SPH_IDX_AHEAD 2
#endif
#ifndef SPH_IDX_AHEAD_F
#define SPH_IDX_AHEAD_F SPH_IDX_AHEAD
#endif

namespace {

constexpr int kPPBlock = SPH_PP_BLOCK;


// Particles [first, first + count).  Slots >= *n_valid (the unused part of the fixed-capacity ghost region of the
// slab decomposition) go to the spare cell g.ncode, which no pass visits.  With index lists given, the particles of
// local x layers 1 and ncl[0] - 2 (the slab's boundary layers) are appended to them, in arrival order.
struct HaloLists {
    int32_t *idx_left, *idx_right;
    uint32_t cap;
};

__global__ void __launch_bounds__(kBlock)
bin_kernel(const __grid_constant__ sph_grid g, const double *__restrict__ r, int first, int count,
           const int32_t *__restrict__ n_valid, int n_owned, uint32_t *__restrict__ cell_count,
           uint32_t *__restrict__ code, uint32_t *__restrict__ rank, sph_status *__restrict__ status, HaloLists hl)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = first + k;
    uint32_t flags = 0;
    bool left = false, right = false;
    if (k < count) {
        if (n_valid && i >= *n_valid) {
            code[i] = g.ncode;
            rank[i] = atomicAdd(cell_count + g.ncode, 1u);
        } else {
            const CellLoc c = locate(g, r[3 * (size_t)i], r[3 * (size_t)i + 1], r[3 * (size_t)i + 2]);
            flags = c.flags;
            code[i] = c.code;
            rank[i] = atomicAdd(cell_count + c.code, 1u);
            left = c.cx == 1;
            right = c.cx == g.ncl[0] - 2;
            // slab decomposition: owned particles live in the inner layers, ghosts in the two outer ones
            if (n_owned > 0 && !g.wrap[0] && ((c.cx == 0 || c.cx == g.ncl[0] - 1) != (i >= n_owned)))
                flags |= SPH_F_OUT_OF_SLAB;
        }
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (flags && (threadIdx.x & 31) == 0) atomicOr(&status->flags, flags);
    if (hl.idx_left) {
        const uint32_t ml = __ballot_sync(0xffffffffu, left), mr = __ballot_sync(0xffffffffu, right);
        const int lane = threadIdx.x & 31;
        uint32_t bl = 0, br = 0;
        if (lane == 0) {
            if (ml) bl = atomicAdd(&status->halo_count[0], (uint32_t)__popc(ml));
            if (mr) br = atomicAdd(&status->halo_count[1], (uint32_t)__popc(mr));
        }
        bl = __shfl_sync(0xffffffffu, bl, 0) + __popc(ml & ((1u << lane) - 1u));
        br = __shfl_sync(0xffffffffu, br, 0) + __popc(mr & ((1u << lane) - 1u));
        if (left && bl < hl.cap) hl.idx_left[bl] = i;
        if (right && br < hl.cap) hl.idx_right[br] = i;
    }
}

__global__ void __launch_bounds__(kBlock)
scatter_kernel(int n, const uint32_t *__restrict__ code, const uint32_t *__restrict__ rank,
               const uint32_t *__restrict__ cell_start, int32_t *__restrict__ perm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[cell_start[code[i]] + rank[i]] = i;
}

// Arrival order inside a cell comes from atomics; make it canonical (ascending original
// index; ghosts of the slab decomposition, whose slots depend on the neighbour rank's packing order,
// by their global id) so that every later sum runs in a fixed order.
__device__ __forceinline__ long long sort_key_of(int32_t v, int n_owned, const int64_t *__restrict__ key)
{
    return (key && n_owned > 0 && v >= n_owned) ? (1ll << 62) + (long long)key[v] : (long long)v;
}

__global__ void __launch_bounds__(kBlock)
cell_sort_kernel(uint32_t ncode, const uint32_t *__restrict__ cell_start, int32_t *__restrict__ perm, int n_owned,
                 const int64_t *__restrict__ key)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncode) return;
    const uint32_t s = cell_start[c], e = cell_start[c + 1];
    for (uint32_t k = s + 1; k < e; ++k) {
        const int32_t v = perm[k];
        const long long kv = sort_key_of(v, n_owned, key);
        uint32_t q = k;
        while (q > s && sort_key_of(perm[q - 1], n_owned, key) > kv) { perm[q] = perm[q - 1]; --q; }
        perm[q] = v;
    }
}

// ------------------------------------------------------------------ exclusive scan (uint32)
constexpr int kScanItems = 8;
constexpr int kScanTile = kBlock * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t &total)
{
    __shared__ uint32_t warp_sums[kBlock / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < kBlock / 32 ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < kBlock / 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < kBlock / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    total = warp_sums[kBlock / 32 - 1];
    const uint32_t base = wid ? warp_sums[wid - 1] : 0u;
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(kBlock)
scan_tile_sums(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ tile_sum)
{
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)k * kBlock + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t total;
    block_exclusive_scan(s, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kBlock)
scan_tile_offsets(uint32_t *__restrict__ tile_sum, int64_t ntiles)
{
    uint32_t carry = 0;
    for (int64_t base = 0; base < ntiles; base += kBlock) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < ntiles ? tile_sum[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, total);
        if (i < ntiles) tile_sum[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sum[ntiles] = carry;
}

__global__ void __launch_bounds__(kBlock)
scan_tile_apply(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int64_t n,
                const uint32_t *__restrict__ tile_off, int64_t ntiles)
{
    // thread t owns kScanItems consecutive elements of the tile
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, total) + tile_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == ntiles - 1 && threadIdx.x == 0) out[n] = tile_off[ntiles];
}

// ------------------------------------------------------------------ gather into Morton order
__global__ void __launch_bounds__(kBlock)
gather_kernel(const __grid_constant__ sph_grid g, int n, const int32_t *__restrict__ perm,
              const double *__restrict__ r, const double *__restrict__ v,
              const double *__restrict__ m, double *__restrict__ pos4, double *__restrict__ vel4,
              float *__restrict__ rel4, int32_t *__restrict__ cnt, const uint32_t *__restrict__ cell_start)
{
    const uint32_t n_live = cell_start[g.ncode];       // particles in real cells
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    if ((uint32_t)a >= n_live) {                       // unused ghost slot: no position, no neighbours
        store4(pos4 + 4 * (size_t)a, 0.0, 0.0, 0.0, 0.0);
        store4(vel4 + 4 * (size_t)a, 0.0, 0.0, 0.0, 0.0);
        reinterpret_cast<float4 *>(rel4)[a] = make_float4(0.f, 0.f, 0.f, 0.f);
        cnt[a] = 0;
        return;
    }
    const size_t i = (size_t)perm[a];
    const double x = r[3 * i], y = r[3 * i + 1], z = r[3 * i + 2];
    const CellLoc c = locate(g, x, y, z);
    store4(pos4 + 4 * (size_t)a, x, y, z, m[i]);
    store4(vel4 + 4 * (size_t)a, v[3 * i], v[3 * i + 1], v[3 * i + 2], 0.0);
    // flags: bit 0 = interior cell (no minimum-image shift can occur), bit 1 = one of the two boundary cell layers
    // of a slab (local x layers 1 and ncl[0] - 2 of a restricted grid: the particles whose neighbours include ghosts)
    const uint32_t edge = (!g.wrap[0] && (c.cx == 1 || c.cx == g.ncl[0] - 2)) ? 2u : 0u;
    reinterpret_cast<float4 *>(rel4)[a] = make_float4(c.rx, c.ry, c.rz, __uint_as_float((c.interior ? 1u : 0u) | edge));
}

// ------------------------------------------------------------------ neighbour pass
__device__ __forceinline__ float wrap32(float d, float L, float half)
{
    if (d > half) d -= L;
    if (d < -half) d += L;
    return d;
}

constexpr int kNlIB = 4;             // particles of the cell tested per staged-candidate load

// One warp per cell.  Lanes 0..26 look up the 27 surrounding cells; their particles (fp32
// cell-relative positions, already shifted into this cell's frame) are staged in shared
// memory; then every lane tests one staged candidate against kNlIB particles of the cell per
// pass.  rsq32 < thr_in accepts, rsq32 >= thr_out rejects, the band in between (and every
// candidate when a position lies outside [-L/4, 5L/4]) is decided by pair_exact.  Accepted
// neighbours are compacted with a ballot into the particle's warp-transposed ELL row.
template <bool SMALL>
__global__ void __launch_bounds__(kNlWarps * 32)
nlist_kernel(const __grid_constant__ sph_grid g, int n, int K,
             const uint32_t *__restrict__ cell_start, const float *__restrict__ rel4,
             const double *__restrict__ pos4, int32_t *__restrict__ nbr, int32_t *__restrict__ cnt,
             sph_status *__restrict__ status, int only_fallback, const int32_t *__restrict__ perm, int n_owned)
{
    extern __shared__ float4 smem_cand[];
    // behind the tile kernels this pass only runs when they gave up (sph_tiles.cu)
    if (only_fallback && !(status->flags & SPH_F_TILE_FALLBACK)) return;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    float4 *S = smem_cand + wib * kNlWin;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool fast = !(status->flags & (SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE));
    // slow mode: nothing is accepted in fp32 and everything is "in the band"
    const float thr_in = fast ? g.thr_in : -1.0f;
    const float thr_out = fast ? g.thr_out : INFINITY;
    const float4 *rel = reinterpret_cast<const float4 *>(rel4);
    const float Lx = (float)g.box[0], Ly = (float)g.box[1], Lz = (float)g.box[2];
    const bool small_x = SMALL && g.ncl[0] < 3, small_y = SMALL && g.ncl[1] < 3, small_z = SMALL && g.ncl[2] < 3;
    uint32_t local_max = 0;

    const uint32_t nwarps = gridDim.x * kNlWarps;
    for (uint32_t c = blockIdx.x * kNlWarps + wib; c < g.ncode; c += nwarps) {
        const uint32_t cs = cell_start[c], ce = cell_start[c + 1];
        if (cs == ce) continue;
        if (n_owned > 0) {                                 // a cell of ghosts: no rows are built for them
            bool ghosts = true;
            for (uint32_t k = cs + lane; k < ce; k += 32) ghosts = ghosts && perm[k] >= n_owned;
            if (__all_sync(0xffffffffu, ghosts)) {
                for (uint32_t k = cs + lane; k < ce; k += 32) cnt[k] = 0;
                continue;
            }
        }
        uint32_t nstart = 0, ncount = 0;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        if (lane < 27) {
            int o[3] = {lane % 3 - 1, (lane / 3) % 3 - 1, lane / 9 - 1};
            bool ok = true;
            uint32_t code = 0;
            if (!SMALL) {
                // neighbour cell codes by arithmetic on the code itself.  Inside a block: dilated-integer
                // +1 is ((part | ~mask) + 1) & mask, -1 is (part - 1) & mask.  Across a block face the
                // block index moves by its stride; across a box face it wraps (or the cell does not exist).
                uint32_t loc = 0, blk = c >> g.lbits;
                uint32_t bc[3];
                block_coords(g, blk, bc[0], bc[1], bc[2]);
                const uint32_t stride[3] = {1u, g.nblk[0], g.nblk[0] * g.nblk[1]};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const uint32_t m = g.mask[d], part = c & m;
                    const bool last_blk = bc[d] == g.nblk[d] - 1u;
                    uint32_t np = part;
                    if (o[d] > 0) {
                        if (last_blk && part == g.top[d]) {            // last layer of the grid
                            np = 0;
                            blk -= bc[d] * stride[d];
                            ok = ok && g.wrap[d];
                        } else if (part == m) {                        // last layer of the block
                            np = 0;
                            blk += stride[d];
                        } else {
                            np = ((part | ~m) + 1u) & m;
                        }
                    } else if (o[d] < 0) {
                        if (bc[d] == 0u && part == 0u) {               // first layer of the grid
                            np = g.top[d];
                            blk += (g.nblk[d] - 1u) * stride[d];
                            ok = ok && g.wrap[d];
                        } else if (part == 0u) {                       // first layer of the block
                            np = m;
                            blk -= stride[d];
                        } else {
                            np = (part - 1u) & m;
                        }
                    }
                    loc |= np;
                }
                code = (blk << g.lbits) | loc;
            } else {
                int cc[3];
                cell_coords(g, c, cc[0], cc[1], cc[2]);
                int nb[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (g.ncl[d] >= 3) {
                        int t = cc[d] + o[d];
                        if (t < 0 || t >= g.ncl[d]) {
                            if (g.wrap[d]) t += (t < 0) ? g.ncl[d] : -g.ncl[d];
                            else ok = false;
                        }
                        nb[d] = t;
                    } else {                          // 1 or 2 layers: visit each layer once
                        const int t = o[d] + 1;
                        if (t >= g.ncl[d]) ok = false;
                        nb[d] = t;
                        o[d] = t - cc[d];
                    }
                }
                code = cell_code(g, nb[0], nb[1], nb[2]);
            }
            if (ok) {
                nstart = cell_start[code];
                ncount = cell_start[code + 1] - nstart;
                fx = (float)((double)o[0] * g.w[0]);
                fy = (float)((double)o[1] * g.w[1]);
                fz = (float)((double)o[2] * g.w[2]);
            }
        }
        uint32_t incl = ncount;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t prefix = incl - ncount;
        const uint32_t maxc = __reduce_max_sync(0xffffffffu, ncount);

        for (uint32_t wbase = 0; wbase < total; wbase += kNlWin) {
            __syncwarp();
            // every lane copies the particles of the neighbour cell it owns
            if (total <= (uint32_t)kNlWin) {                    // the usual case: one window
                const float4 *src = rel + nstart;
                float4 *dst = S + prefix;
                for (uint32_t t = 0; t < maxc; ++t) {
                    if (t < ncount) {
                        const float4 p = __ldg(src + t);
                        dst[t] = make_float4(p.x + fx, p.y + fy, p.z + fz, __int_as_float((int)(nstart + t)));
                    }
                }
            } else {
                for (uint32_t t = 0; t < maxc; ++t) {
                    const uint32_t s = prefix + t - wbase;      // wraps below the window: fails s < kNlWin
                    if (t < ncount && s < (uint32_t)kNlWin) {
                        const float4 p = __ldg(rel + nstart + t);
                        S[s] = make_float4(p.x + fx, p.y + fy, p.z + fz, __int_as_float((int)(nstart + t)));
                    }
                }
            }
            const uint32_t nS = min((uint32_t)kNlWin, total - wbase);
            const uint32_t nSp = (nS + 31u) & ~31u;
            // pad the last chunk with candidates at infinity: rsq = +inf fails every threshold
            if (nS + lane < nSp) S[nS + lane] = make_float4(1.0e30f, 1.0e30f, 1.0e30f, __int_as_float(-1));
            __syncwarp();
            for (uint32_t a0 = cs; a0 < ce; a0 += kNlIB) {
                float px[kNlIB], py[kNlIB], pz[kNlIB], tin[kNlIB], tout[kNlIB];
                uint32_t cnt_i[kNlIB];
                int self[kNlIB];
                int32_t *row[kNlIB];
#pragma unroll
                for (int u = 0; u < kNlIB; ++u) {
                    const bool act = a0 + u < ce;
                    const uint32_t a = act ? a0 + u : ce - 1;
                    const float4 pi = __ldg(rel + a);
                    px[u] = pi.x; py[u] = pi.y; pz[u] = pi.z;
                    tin[u] = act ? thr_in : -1.0f;                 // padding rows accept nothing ...
                    tout[u] = act ? thr_out : -1.0f;               // ... and have no band
                    self[u] = (int)a;
                    cnt_i[u] = (wbase && act) ? (uint32_t)cnt[a] : 0u;
                    row[u] = nbr + ((size_t)(a >> 5) * (size_t)K) * 32 + (a & 31);
                }
                for (uint32_t s0 = 0; s0 < nSp; s0 += 32) {
                    const float4 cd = S[s0 + lane];
                    const int j = __float_as_int(cd.w);
                    bool in[kNlIB], band = false;
#pragma unroll
                    for (int u = 0; u < kNlIB; ++u) {
                        float dx = cd.x - px[u], dy = cd.y - py[u], dz = cd.z - pz[u];
                        if (SMALL) {
                            if (small_x) dx = wrap32(dx, Lx, 0.5f * Lx);
                            if (small_y) dy = wrap32(dy, Ly, 0.5f * Ly);
                            if (small_z) dz = wrap32(dz, Lz, 0.5f * Lz);
                        }
                        const float rsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        in[u] = rsq < tin[u] && j != self[u];
                        band = band || (!in[u] && rsq < tout[u] && j != self[u]);
                    }
                    if (__any_sync(0xffffffffu, band)) {           // rare: decide in fp64
#pragma unroll
                        for (int u = 0; u < kNlIB; ++u) {
                            float dx = cd.x - px[u], dy = cd.y - py[u], dz = cd.z - pz[u];
                            if (SMALL) {
                                if (small_x) dx = wrap32(dx, Lx, 0.5f * Lx);
                                if (small_y) dy = wrap32(dy, Ly, 0.5f * Ly);
                                if (small_z) dz = wrap32(dz, Lz, 0.5f * Lz);
                            }
                            const float rsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            if (!in[u] && rsq < tout[u] && j != self[u] && j >= 0)
                                in[u] = pair_exact(g, pos4, self[u], j);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kNlIB; ++u) {
                        const uint32_t mask = __ballot_sync(0xffffffffu, in[u]);
                        const uint32_t pos = cnt_i[u] + __popc(mask & lt_mask);
                        if (in[u] && pos < (uint32_t)K) row[u][(size_t)pos * 32] = j;
                        cnt_i[u] += __popc(mask);
                    }
                }
#pragma unroll
                for (int u = 0; u < kNlIB; ++u) {
                    if (a0 + u < ce) {
                        if (lane == 0) cnt[a0 + u] = (int32_t)cnt_i[u];
                        local_max = max(local_max, cnt_i[u]);
                    }
                }
            }
        }
    }
    if (lane == 0 && local_max > 0) {
        if (local_max > *(volatile uint32_t *)&status->max_count) atomicMax(&status->max_count, local_max);
        if (local_max > (uint32_t)K) atomicOr(&status->flags, SPH_F_NBR_OVERFLOW);
    }
}

// particles that sit in no cell the neighbour pass visited cannot exist (every particle is
// in a cell), but rows of EMPTY trailing lanes must read as zero neighbours.
// ------------------------------------------------------------------ per-particle passes
struct Interior {
    bool skip;   // warp-uniform: no pair of this warp can need the minimum-image shift
};

__device__ __forceinline__ bool cell_is_interior(const sph_grid &, uint32_t flag)
{
    return (flag & 1u) != 0u;   // computed once per particle by gather_kernel (rel4[., 3]; bit 1: slab boundary layer)
}

constexpr int kRowU = SPH_ROW_U;             // neighbours gathered per pipeline stage

template <bool UNIFORM_H, bool WRAP>
__device__ __forceinline__ double density_row(const sph_grid &g, const double *__restrict__ pos4,
                                              const int32_t *__restrict__ perm,
                                              const double *__restrict__ h_orig,
                                              const int32_t *__restrict__ row, size_t stride, int count,
                                              int orig, int self, double ax, double ay, double az, double hinv,
                                              double qn)
{
    double acc = 0.0;
    const double hx = g.box[0] / 2., hy = g.box[1] / 2., hz = g.box[2] / 2.;
    // indices are fetched one stage ahead, so that kRowU gathers are in flight while the
    // previous kRowU pairs are evaluated
    int jn[SPH_IDX_AHEAD][kRowU];          // indices of the next SPH_IDX_AHEAD stages
#pragma unroll
    for (int st = 0; st < SPH_IDX_AHEAD; ++st)
#pragma unroll
        for (int u = 0; u < kRowU; ++u)
            jn[st][u] = st * kRowU + u < count ? row[(size_t)(st * kRowU + u) * stride] : self;
    for (int k0 = 0; k0 < count; k0 += kRowU) {
        double bx[kRowU], by[kRowU], bz[kRowU], bm[kRowU];
        int j[kRowU];
#pragma unroll
        for (int u = 0; u < kRowU; ++u) {
            j[u] = jn[0][u];
            load4(pos4 + 4 * (size_t)j[u], bx[u], by[u], bz[u], bm[u]);
        }
#pragma unroll
        for (int u = 0; u < kRowU; ++u) {
            const int kn = k0 + SPH_IDX_AHEAD * kRowU + u;
#pragma unroll
            for (int st = 0; st + 1 < SPH_IDX_AHEAD; ++st) jn[st][u] = jn[st + 1][u];
            jn[SPH_IDX_AHEAD - 1][u] = kn < count ? row[(size_t)kn * stride] : self;
        }
#pragma unroll
        for (int u = 0; u < kRowU; ++u) {
            double dx = bx[u] - ax, dy = by[u] - ay, dz = bz[u] - az;
            if (WRAP) {
                dx = min_image(dx, g.box[0], hx);
                dy = min_image(dy, g.box[1], hy);
                dz = min_image(dz, g.box[2], hz);
            }
            const double rsq = rsq_exact(dx, dy, dz);
            const double rr = sqrt(rsq);
            double hi = hinv, q = qn;
            if (!UNIFORM_H) {
                const int oj = perm[j[u]];
                const double h = h_orig[oj < orig ? oj : orig];   // properties.py:88: h of the first member
                hi = 1.0 / h;
                q = lucy_norm3(h);
            }
            const double s = rr * hi;
            if (s < 1.0 && k0 + u < count) {                      // spkernel.py:106
                const double t = 1.0 - s;
                acc += (q * (1.0 + 3.0 * s) * (t * t * t)) * bm[u];   // spkernel.py:107, properties.py:90-91
            }
        }
    }
    return acc;
}

// LPP lanes cooperate on one particle: lane q of the group takes neighbours q, q+LPP, ... of the
// row, so that the lanes of a group gather consecutive rows (neighbours from one cell are
// contiguous in the sorted arrays) and the per-lane trip counts even out; the partial sums are
// combined with shuffles.
template <bool UNIFORM_H, int LPP>
__global__ void __launch_bounds__(kPPBlock, SPH_DENS_MINB)
density_kernel(const __grid_constant__ sph_grid g, int n, int K, const double *__restrict__ pos4,
               double *__restrict__ vel4, const float *__restrict__ rel4,
               const int32_t *__restrict__ perm, const int32_t *__restrict__ nbr,
               const int32_t *__restrict__ cnt, const sph_status *__restrict__ status,
               const double *__restrict__ h_orig, sph_eos eos, int list_fresh, int long_range, int n_owned,
               double *__restrict__ rho_out, double *__restrict__ p_out, double *__restrict__ pco_out,
               double *__restrict__ u_out, double *__restrict__ t_io)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = gt / LPP, q = gt % LPP;
    const bool active = a < n;
    double ax = 0, ay = 0, az = 0, am = 0;
    int count = 0, orig = 0;
    bool interior = true;
    if (active) {
        load4(pos4 + 4 * (size_t)a, ax, ay, az, am);
        orig = perm[a];
        count = (n_owned > 0 && orig >= n_owned) ? 0 : min(cnt[a], K);       // nothing is computed for ghosts
        interior = cell_is_interior(g, __float_as_uint(reinterpret_cast<const float4 *>(rel4)[a].w));
    }
    const bool can_skip = list_fresh && !(status->flags & (SPH_F_OUT_OF_BOX | SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE));
    const bool skip = __all_sync(0xffffffffu, interior) && can_skip;
    const double h0 = h_orig[0];
    const double hinv = 1.0 / h0, qn = lucy_norm3(h0);
    const int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K + q) * 32 + (a & 31);
    const int mine = count > q ? (count - q + LPP - 1) / LPP : 0;
    double sum;
    if (skip) sum = density_row<UNIFORM_H, false>(g, pos4, perm, h_orig, row, (size_t)32 * LPP, mine, orig, a, ax, ay, az, hinv, qn);
    else sum = density_row<UNIFORM_H, true>(g, pos4, perm, h_orig, row, (size_t)32 * LPP, mine, orig, a, ax, ay, az, hinv, qn);
#pragma unroll
    for (int o = 1; o < LPP; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (!active || q != 0 || (n_owned > 0 && orig >= n_owned)) return;
    // properties.py:76-77: every particle starts from W(0; h[0]) -- not m_i * W(0; h_i)
    const double rho = qn + sum;
    rho_out[orig] = rho;
    if (long_range == 1) return;
    double t_eos, t_new;
    if (long_range == 2) {
        // SpamComplete: u is the integrated state and T follows from it (spam_complete_force.py:134,151-152:
        // calc_vdw_temp, then T[T < 0] = 0); u is left alone
        t_new = (u_out[orig] + eos.adash * rho) / eos.kbdash;               // properties.py:49
        if (t_new < 0.0) t_new = 0.0;
        t_eos = t_new;
    } else {
        t_eos = t_io[orig];
        const double u = t_eos * eos.kbdash - eos.adash * rho;              // properties.py:46,119
        u_out[orig] = u;
        t_new = (u + eos.adash * rho) / eos.kbdash;                         // properties.py:49,120
    }
    const double p = (rho * eos.kbdash * t_eos) / (1 - rho * eos.bdash);    // properties.py:41
    const double pco = -eos.adash * rho * rho;
    p_out[orig] = p;
    pco_out[orig] = pco;
    t_io[orig] = t_new;
    vel4[4 * (size_t)a + 3] = p / (rho * rho);                              // forces.py:353 operand
}

__global__ void __launch_bounds__(kBlock)
pressure_term_kernel(int n, int first_orig, const int32_t *__restrict__ perm, const double *__restrict__ press,
                     const double *__restrict__ rho, double *__restrict__ vel4)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int o = perm[a];
    if (o < first_orig) return;
    const double d = rho[o];
    vel4[4 * (size_t)a + 3] = press[o] / (d * d);
}

struct ForceAcc { double ax, ay, az, du; };

constexpr int kRowUF = SPH_ROW_UF;   // force: 8 doubles per neighbour, so a shorter stage
constexpr int kRowUC = 2;            // conduction (not register-capped)

template <bool UNIFORM_H, bool WRAP>
__device__ __forceinline__ ForceAcc force_row(const sph_grid &g, const double *__restrict__ pos4,
                                              const double *__restrict__ vel4,
                                              const int32_t *__restrict__ perm,
                                              const double *__restrict__ h_orig,
                                              const int32_t *__restrict__ row, size_t stride, int count,
                                              int orig, int self, double px, double py, double pz, double vx, double vy,
                                              double vz, double Ai, double hinv, double c2,
                                              double fcutsq, bool two_d)
{
    ForceAcc f = {0.0, 0.0, 0.0, 0.0};
    const double hx = g.box[0] / 2., hy = g.box[1] / 2., hz = g.box[2] / 2.;
    int jn[SPH_IDX_AHEAD_F][kRowUF];       // indices of the next SPH_IDX_AHEAD_F stages
#pragma unroll
    for (int st = 0; st < SPH_IDX_AHEAD_F; ++st)
#pragma unroll
        for (int u = 0; u < kRowUF; ++u)
            jn[st][u] = st * kRowUF + u < count ? row[(size_t)(st * kRowUF + u) * stride] : self;
    for (int k0 = 0; k0 < count; k0 += kRowUF) {
        double bx[kRowUF], by[kRowUF], bz[kRowUF], bm[kRowUF], wx[kRowUF], wy[kRowUF], wz[kRowUF], Aj[kRowUF];
        int j[kRowUF];
#pragma unroll
        for (int u = 0; u < kRowUF; ++u) {
            j[u] = jn[0][u];
            load4(pos4 + 4 * (size_t)j[u], bx[u], by[u], bz[u], bm[u]);
            load4(vel4 + 4 * (size_t)j[u], wx[u], wy[u], wz[u], Aj[u]);
        }
#pragma unroll
        for (int u = 0; u < kRowUF; ++u) {
            const int kn = k0 + SPH_IDX_AHEAD_F * kRowUF + u;
#pragma unroll
            for (int st = 0; st + 1 < SPH_IDX_AHEAD_F; ++st) jn[st][u] = jn[st + 1][u];
            jn[SPH_IDX_AHEAD_F - 1][u] = kn < count ? row[(size_t)kn * stride] : self;
        }
#pragma unroll
        for (int u = 0; u < kRowUF; ++u) {
            // Pair (i<j in original order) contributes +a to i and -a to j with dr = r_j - r_i.
            // Seen from this particle: dr = r_other - r_self and the term enters with +.
            double dx = bx[u] - px, dy = by[u] - py, dz = bz[u] - pz;
            if (WRAP) {
                dx = min_image(dx, g.box[0], hx);
                dy = min_image(dy, g.box[1], hy);
                dz = min_image(dz, g.box[2], hz);
            }
            const double rsq = rsq_exact(dx, dy, dz);
            const double rr = sqrt(rsq);
            double hi = hinv, cc = c2;
            if (!UNIFORM_H) {
                const int oj = perm[j[u]];
                const double h = h_orig[oj < orig ? oj : orig];
                hi = 1.0 / h;
                cc = -12.0 * lucy_norm3(h) * hi * hi;
            }
            const double s = rr * hi;
            // forces.py:40 cutoff on rij^2; spkernel.py:106,109: zero outside h and at r == 0
            if (s < 1.0 && rr * rr <= fcutsq && k0 + u < count) {
                // q(-12 r^3/h^4 + 24 r^2/h^3 - 12 r/h^2)/r = -(12 q / h^2) (1 - r/h)^2   (spkernel.py:113-114)
                const double t = 1.0 - s;
                const double fac = (cc * (t * t)) * (Ai + Aj[u]);   // ps * dW/dr / r  (forces.py:353-357)
                const double gx = fac * dx, gy = fac * dy, gz = two_d ? 0.0 : fac * dz;
                f.ax += gx;
                f.ay += gy;
                f.az += gz;
                // du = 0.5 * a . dv with dv = v_j - v_i; symmetric in the pair (forces.py:366-368)
                const double dot = gx * (wx[u] - vx) + gy * (wy[u] - vy) + gz * (wz[u] - vz);
                f.du += (0.5 * dot) * bm[u];
            }
        }
    }
    return f;
}

template <bool UNIFORM_H, int LPP>
__global__ void __launch_bounds__(kPPBlock, SPH_FORCE_MINB)
force_kernel(const __grid_constant__ sph_grid g, int n, int K, const double *__restrict__ pos4,
             const double *__restrict__ vel4, const float *__restrict__ rel4,
             const int32_t *__restrict__ perm, const int32_t *__restrict__ nbr,
             const int32_t *__restrict__ cnt, const sph_status *__restrict__ status,
             const double *__restrict__ h_orig, int list_fresh, double fcutsq, int dim, int n_owned, int store,
             int part, double *__restrict__ vdot, double *__restrict__ udot)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = gt / LPP, q = gt % LPP;
    const bool active = a < n;
    double px = 0, py = 0, pz = 0, pm = 0, vx = 0, vy = 0, vz = 0, Ai = 0;
    int count = 0, orig = 0;
    bool interior = true, in_part = true;
    if (active) {
        load4(pos4 + 4 * (size_t)a, px, py, pz, pm);
        load4(vel4 + 4 * (size_t)a, vx, vy, vz, Ai);
        orig = perm[a];
        count = (n_owned > 0 && orig >= n_owned) ? 0 : min(cnt[a], K);       // nothing is computed for ghosts
        const uint32_t fl = __float_as_uint(reinterpret_cast<const float4 *>(rel4)[a].w);
        interior = cell_is_interior(g, fl);
        // part 1: every particle but those of the slab's boundary layers, part 2: only those (0: all)
        in_part = part == 0 || ((fl & 2u) != 0u) == (part == 2);
        if (!in_part) count = 0;
    }
    const bool can_skip = list_fresh && !(status->flags & (SPH_F_OUT_OF_BOX | SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE));
    const bool skip = __all_sync(0xffffffffu, interior) && can_skip;
    const double h0 = h_orig[0];
    const double hinv = 1.0 / h0, c2 = -12.0 * lucy_norm3(h0) * hinv * hinv;
    const int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K + q) * 32 + (a & 31);
    const int mine = count > q ? (count - q + LPP - 1) / LPP : 0;
    ForceAcc f;
    if (skip) f = force_row<UNIFORM_H, false>(g, pos4, vel4, perm, h_orig, row, (size_t)32 * LPP, mine, orig, a, px, py, pz, vx, vy, vz, Ai, hinv, c2, fcutsq, dim == 2);
    else f = force_row<UNIFORM_H, true>(g, pos4, vel4, perm, h_orig, row, (size_t)32 * LPP, mine, orig, a, px, py, pz, vx, vy, vz, Ai, hinv, c2, fcutsq, dim == 2);
#pragma unroll
    for (int o = 1; o < LPP; o <<= 1) {
        f.ax += __shfl_xor_sync(0xffffffffu, f.ax, o);
        f.ay += __shfl_xor_sync(0xffffffffu, f.ay, o);
        f.az += __shfl_xor_sync(0xffffffffu, f.az, o);
        f.du += __shfl_xor_sync(0xffffffffu, f.du, o);
    }
    if (!active || !in_part || q != 0 || (n_owned > 0 && orig >= n_owned)) return;
    (void)pm;
    // the reference accumulates into vdot/udot (particles.py:549-550 zeroes them per evaluation); `store` says this
    // is the first force after that zeroing, so the fill and the read-modify-write are both saved (0 + x == x)
    if (store) {
        vdot[3 * (size_t)orig] = f.ax;
        vdot[3 * (size_t)orig + 1] = f.ay;
        vdot[3 * (size_t)orig + 2] = f.az;
        udot[orig] = f.du;
    } else {
        vdot[3 * (size_t)orig] += f.ax;
        vdot[3 * (size_t)orig + 1] += f.ay;
        vdot[3 * (size_t)orig + 2] += f.az;
        udot[orig] += f.du;
    }
}

// The force pass over the particles of a slab's two boundary cell layers ONLY (part 2 of sph_force on a restricted grid):
// 16 lanes per cell of local x layers 1 and ncl[0] - 2, found through the cell table, so that the launch costs what
// those 1-2 % of the particles cost -- the flag test of force_kernel walks every warp of the system for them.  This is the
// launch that waits for the ghosts' (p, rho); everything else runs while they travel (SlabSphEvaluator.overlap_b).
template <bool UNIFORM_H>
__global__ void __launch_bounds__(kPPBlock, SPH_FORCE_MINB)
force_edge_kernel(const __grid_constant__ sph_grid g, int K, const double *__restrict__ pos4,
                  const double *__restrict__ vel4, const float *__restrict__ rel4,
                  const int32_t *__restrict__ perm, const int32_t *__restrict__ nbr,
                  const int32_t *__restrict__ cnt, const uint32_t *__restrict__ cell_start,
                  const sph_status *__restrict__ status, const double *__restrict__ h_orig, int list_fresh,
                  double fcutsq, int dim, int n_owned, int store, double *__restrict__ vdot, double *__restrict__ udot)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = gt >> 4, sub = gt & 15;
    const int ny = g.ncl[1], per = ny * g.ncl[2];
    const int sides = g.ncl[0] - 2 > 1 ? 2 : 1;
    uint32_t first = 0, here = 0;
    if (c < sides * per) {
        const int side = c >= per ? 1 : 0, rem = c - side * per;
        const uint32_t code = cell_code(g, side ? g.ncl[0] - 2 : 1, rem % ny, rem / ny);
        first = cell_start[code];
        here = cell_start[code + 1] - first;
    }
    const int trips = (int)__reduce_max_sync(0xffffffffu, (here + 15u) >> 4);
    const bool can_skip = list_fresh && !(status->flags & (SPH_F_OUT_OF_BOX | SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE));
    const double h0 = h_orig[0];
    const double hinv = 1.0 / h0, c2 = -12.0 * lucy_norm3(h0) * hinv * hinv;
    for (int tr = 0; tr < trips; ++tr) {
        const uint32_t k = (uint32_t)sub + 16u * (uint32_t)tr;
        const bool active = k < here;
        const int a = active ? (int)(first + k) : 0;
        double px = 0, py = 0, pz = 0, pm = 0, vx = 0, vy = 0, vz = 0, Ai = 0;
        int count = 0, orig = 0;
        bool interior = true;
        if (active) {
            load4(pos4 + 4 * (size_t)a, px, py, pz, pm);
            load4(vel4 + 4 * (size_t)a, vx, vy, vz, Ai);
            orig = perm[a];
            count = orig >= n_owned ? 0 : min(cnt[a], K);
            interior = cell_is_interior(g, __float_as_uint(reinterpret_cast<const float4 *>(rel4)[a].w));
        }
        const bool skip = __all_sync(0xffffffffu, interior) && can_skip;
        const int32_t *row = nbr + (size_t)(a >> 5) * (size_t)K * 32 + (a & 31);
        ForceAcc f;
        if (skip) f = force_row<UNIFORM_H, false>(g, pos4, vel4, perm, h_orig, row, (size_t)32, count, orig, a, px, py, pz, vx, vy, vz, Ai, hinv, c2, fcutsq, dim == 2);
        else f = force_row<UNIFORM_H, true>(g, pos4, vel4, perm, h_orig, row, (size_t)32, count, orig, a, px, py, pz, vx, vy, vz, Ai, hinv, c2, fcutsq, dim == 2);
        (void)pm;
        if (!active || orig >= n_owned) continue;
        if (store) {
            vdot[3 * (size_t)orig] = f.ax;
            vdot[3 * (size_t)orig + 1] = f.ay;
            vdot[3 * (size_t)orig + 2] = f.az;
            udot[orig] = f.du;
        } else {
            vdot[3 * (size_t)orig] += f.ax;
            vdot[3 * (size_t)orig + 1] += f.ay;
            vdot[3 * (size_t)orig + 2] += f.az;
            udot[orig] += f.du;
        }
    }
}

// ------------------------------------------------------------------ heat conduction (c_forces.pyx:196-239)
__global__ void __launch_bounds__(kBlock)
flux_term_kernel(int n, const int32_t *__restrict__ perm, const double *__restrict__ jq,
                 const double *__restrict__ rho, double *__restrict__ aux4)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const size_t o = (size_t)perm[a];
    const double d = rho[o], inv = 1.0 / (d * d);
    store4(aux4 + 4 * (size_t)a, jq[3 * o] * inv, jq[3 * o + 1] * inv, jq[3 * o + 2] * inv, 0.0);   // q / rho^2
}

template <bool UNIFORM_H, bool WRAP>
__device__ __forceinline__ double conduction_row(const sph_grid &g, const double *__restrict__ pos4,
                                                 const double *__restrict__ aux4,
                                                 const int32_t *__restrict__ perm,
                                                 const double *__restrict__ h_orig,
                                                 const int32_t *__restrict__ row, int count, int orig, int self,
                                                 double px, double py, double pz, double qx, double qy,
                                                 double qz, double hinv, double c2)
{
    double acc = 0.0;
    const double hx = g.box[0] / 2., hy = g.box[1] / 2., hz = g.box[2] / 2.;
    int jn[kRowUC];
#pragma unroll
    for (int u = 0; u < kRowUC; ++u) jn[u] = u < count ? row[(size_t)u * 32] : self;
    for (int k0 = 0; k0 < count; k0 += kRowUC) {
        double bx[kRowUC], by[kRowUC], bz[kRowUC], bm[kRowUC], ex[kRowUC], ey[kRowUC], ez[kRowUC], ew[kRowUC];
        int j[kRowUC];
#pragma unroll
        for (int u = 0; u < kRowUC; ++u) {
            j[u] = jn[u];
            load4(pos4 + 4 * (size_t)j[u], bx[u], by[u], bz[u], bm[u]);
            load4(aux4 + 4 * (size_t)j[u], ex[u], ey[u], ez[u], ew[u]);
        }
#pragma unroll
        for (int u = 0; u < kRowUC; ++u) {
            const int kn = k0 + kRowUC + u;
            jn[u] = kn < count ? row[(size_t)kn * 32] : self;
        }
#pragma unroll
        for (int u = 0; u < kRowUC; ++u) {
            double dx = bx[u] - px, dy = by[u] - py, dz = bz[u] - pz;
            if (WRAP) {
                dx = min_image(dx, g.box[0], hx);
                dy = min_image(dy, g.box[1], hy);
                dz = min_image(dz, g.box[2], hz);
            }
            const double rr = sqrt(rsq_exact(dx, dy, dz));
            double hi = hinv, cc = c2;
            if (!UNIFORM_H) {
                const int oj = perm[j[u]];
                const double h = h_orig[oj < orig ? oj : orig];
                hi = 1.0 / h;
                cc = -12.0 * lucy_norm3(h) * hi * hi;
            }
            const double s = rr * hi;
            if (s < 1.0 && k0 + u < count) {
                const double t = 1.0 - s;
                const double fac = cc * (t * t);                       // dW/dx_a = fac * dx_a
                // c_forces.pyx:228-237 seen from this particle: udot_self -= sum_a (Q_self + Q_other)_a dW_a m_other
                acc -= (((qx + ex[u]) * (fac * dx) + (qy + ey[u]) * (fac * dy)) + (qz + ez[u]) * (fac * dz)) * bm[u];
            }
        }
    }
    return acc;
}

template <bool UNIFORM_H>
__global__ void __launch_bounds__(kBlock)
conduction_kernel(const __grid_constant__ sph_grid g, int n, int K, const double *__restrict__ pos4,
                  const double *__restrict__ aux4, const float *__restrict__ rel4,
                  const int32_t *__restrict__ perm, const int32_t *__restrict__ nbr,
                  const int32_t *__restrict__ cnt, const sph_status *__restrict__ status,
                  const double *__restrict__ h_orig, int list_fresh, double *__restrict__ udot)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = a < n;
    double px = 0, py = 0, pz = 0, pm = 0, qx = 0, qy = 0, qz = 0, qw = 0;
    int count = 0, orig = 0;
    bool interior = true;
    if (active) {
        load4(pos4 + 4 * (size_t)a, px, py, pz, pm);
        load4(aux4 + 4 * (size_t)a, qx, qy, qz, qw);
        count = min(cnt[a], K);
        orig = perm[a];
        interior = cell_is_interior(g, __float_as_uint(reinterpret_cast<const float4 *>(rel4)[a].w));
    }
    const bool can_skip = list_fresh && !(status->flags & (SPH_F_OUT_OF_BOX | SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE));
    const bool skip = __all_sync(0xffffffffu, interior) && can_skip;
    const double h0 = h_orig[0];
    const double hinv = 1.0 / h0, c2 = -12.0 * lucy_norm3(h0) * hinv * hinv;
    const int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K) * 32 + (a & 31);
    double acc;
    if (skip) acc = conduction_row<UNIFORM_H, false>(g, pos4, aux4, perm, h_orig, row, count, orig, a, px, py, pz, qx, qy, qz, hinv, c2);
    else acc = conduction_row<UNIFORM_H, true>(g, pos4, aux4, perm, h_orig, row, count, orig, a, px, py, pz, qx, qy, qz, hinv, c2);
    (void)pm; (void)qw;
    if (active) udot[orig] += acc;
}

// ------------------------------------------------------------------ list maintenance / export
__global__ void __launch_bounds__(kBlock)
compress_kernel(const __grid_constant__ sph_grid g, int n, int K, const double *__restrict__ pos4,
                int32_t *__restrict__ nbr, int32_t *__restrict__ cnt)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K) * 32 + (a & 31);
    const int count = min(cnt[a], K);
    int q = 0;
    for (int k = 0; k < count; ++k) {
        const int j = row[(size_t)k * 32];
        if (pair_exact(g, pos4, a, j)) row[(size_t)(q++) * 32] = j;
    }
    cnt[a] = q;
}

__global__ void __launch_bounds__(kBlock)
pairs_count_kernel(int n, int K, const int32_t *__restrict__ perm, const int32_t *__restrict__ nbr,
                   const int32_t *__restrict__ cnt, uint32_t *__restrict__ row_count)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K) * 32 + (a & 31);
    const int count = min(cnt[a], K);
    const int oi = perm[a];
    uint32_t c = 0;
    for (int k = 0; k < count; ++k) c += perm[row[(size_t)k * 32]] > oi;
    row_count[oi] = c;
}

constexpr int kExportMax = 256;   // neighbours sorted in local memory per particle

__global__ void __launch_bounds__(128)
pairs_fill_kernel(int n, int K, const int32_t *__restrict__ perm, const int32_t *__restrict__ nbr,
                  const int32_t *__restrict__ cnt, const int64_t *__restrict__ row_start,
                  int32_t *__restrict__ iap, int64_t cap)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K) * 32 + (a & 31);
    const int count = min(cnt[a], K);
    const int oi = perm[a];
    int32_t js[kExportMax];
    int m = 0;
    int64_t base = row_start[oi];
    // rows longer than kExportMax are emitted in sorted chunks by repeated selection
    int last = oi;
    for (;;) {
        m = 0;
        for (int k = 0; k < count; ++k) {
            const int oj = perm[row[(size_t)k * 32]];
            if (oj <= last) continue;
            // keep the kExportMax smallest candidates above `last`, sorted ascending
            int q = m < kExportMax ? m : kExportMax;
            if (q == kExportMax && oj >= js[kExportMax - 1]) continue;
            if (q == kExportMax) q = kExportMax - 1;
            while (q > 0 && js[q - 1] > oj) { js[q] = js[q - 1]; --q; }
            js[q] = oj;
            if (m < kExportMax) ++m;
        }
        for (int t = 0; t < m; ++t) {
            if (base + t < cap) {
                iap[2 * (base + t)] = oi;
                iap[2 * (base + t) + 1] = js[t];
            }
        }
        if (m < kExportMax) break;
        base += m;
        last = js[m - 1];
    }
}

__global__ void __launch_bounds__(kBlock)
separations_kernel(double Lx, double Ly, double Lz, const int32_t *__restrict__ iap, int64_t nip,
                   const double *__restrict__ r, const double *__restrict__ v,
                   double *__restrict__ drij, double *__restrict__ rij, double *__restrict__ rsq,
                   double *__restrict__ dv)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nip) return;
    const size_t i = (size_t)iap[2 * k], j = (size_t)iap[2 * k + 1];
    // neighbour_list.py:66-82
    const double dx = min_image(__dsub_rn(r[3 * j], r[3 * i]), Lx, Lx / 2.);
    const double dy = min_image(__dsub_rn(r[3 * j + 1], r[3 * i + 1]), Ly, Ly / 2.);
    const double dz = min_image(__dsub_rn(r[3 * j + 2], r[3 * i + 2]), Lz, Lz / 2.);
    const double s = rsq_exact(dx, dy, dz);
    drij[3 * k] = dx; drij[3 * k + 1] = dy; drij[3 * k + 2] = dz;
    rsq[k] = s;
    rij[k] = sqrt(s);
    dv[3 * k] = v[3 * j] - v[3 * i];
    dv[3 * k + 1] = v[3 * j + 1] - v[3 * i + 1];
    dv[3 * k + 2] = v[3 * j + 2] - v[3 * i + 2];
}

__global__ void __launch_bounds__(kBlock)
pair_kernels_kernel(const int32_t *__restrict__ iap, int64_t nip, const double *__restrict__ rij,
                    const double *__restrict__ drij, const double *__restrict__ h,
                    double *__restrict__ wij, double *__restrict__ dwij)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nip) return;
    const double hh = h[iap[2 * k]];                       // properties.py:88
    double rr = rij[k];
    if (rr < 0) rr = fabs(rr);                             // spkernel.py:104-105
    const double q = lucy_norm3(hh);
    double w = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
    if (rr < hh) {
        const double t = 1. - rr / hh;
        w = q * (1 + 3. * rr / hh) * (t * t * t);          // spkernel.py:107
        if (rr != 0) {
            const double h2 = hh * hh;
            const double f = q * ((-12. / (h2 * h2)) * (rr * rr * rr) + (24. / (h2 * hh)) * (rr * rr)
                                  - (12. * rr / h2));      // spkernel.py:113-114
            gx = f * drij[3 * k] / rr;
            gy = f * drij[3 * k + 1] / rr;
            gz = f * drij[3 * k + 2] / rr;
        }
    }
    wij[k] = w;
    dwij[3 * k] = gx; dwij[3 * k + 1] = gy; dwij[3 * k + 2] = gz;
}

__global__ void __launch_bounds__(kBlock)
ponder_kernel(const double *__restrict__ r_old, const double *__restrict__ r, int n,
              sph_status *__restrict__ status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (i < n) {
        // neighbour_list.py:230-231
        const double a = r_old[3 * (size_t)i] - r[3 * (size_t)i];
        const double b = r_old[3 * (size_t)i + 1] - r[3 * (size_t)i + 1];
        const double c = r_old[3 * (size_t)i + 2] - r[3 * (size_t)i + 2];
        s = rsq_exact(a, b, c);
        if (!(s >= 0.0)) s = 0.0;                          // NaN never triggers `dsq > tol`
    }
    unsigned long long bits = (unsigned long long)__double_as_longlong(s);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = t > bits ? t : bits;
    }
    if ((threadIdx.x & 31) == 0 && bits > *(volatile unsigned long long *)&status->dsq_max_bits)
        atomicMax(&status->dsq_max_bits, bits);
}

__global__ void ponder_decide_kernel(sph_status *status, double tol_sq)
{
    const double dsq = __longlong_as_double((long long)status->dsq_max_bits);
    status->rebuild = dsq > tol_sq ? 1u : 0u;              // neighbour_list.py:233-234
}

__global__ void status_reset_kernel(sph_status *status)
{
    if (threadIdx.x < sizeof(sph_status) / 4) reinterpret_cast<uint32_t *>(status)[threadIdx.x] = 0u;
}

__global__ void flags_clear_kernel(sph_status *status, uint32_t bits)
{
    status->flags &= ~bits;
}

__global__ void __launch_bounds__(kBlock)
axpy_kernel(double *__restrict__ x, const double *__restrict__ a, const double *__restrict__ b,
            double s, int64_t len)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) x[i] = a[i] + s * b[i];
}

__global__ void __launch_bounds__(kBlock)
box_kernel(double Lx, double Ly, double Lz, int kind, double *__restrict__ r, double *__restrict__ v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double L[3] = {Lx, Ly, Lz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double x = r[3 * (size_t)i + d];
        if (kind == 0) {                                   // box.py:53-73 MirrorBox
            if (x > L[d]) { x = L[d]; v[3 * (size_t)i + d] = -v[3 * (size_t)i + d]; }
            if (x < 0) { x = 0; v[3 * (size_t)i + d] = -v[3 * (size_t)i + d]; }
        } else if (kind == 1) {                            // box.py:35-47 PeriodicBox
            if (x > L[d]) x = 0;
            if (x < 0) x = L[d];
        } else {                                           // true periodic image (not the reference's reset)
            if (x >= L[d] || x < 0) {
                x -= L[d] * floor(x / L[d]);
                if (x >= L[d]) x = 0;                      // -tiny + L rounds to L
            }
        }
        r[3 * (size_t)i + d] = x;
    }
}

// ------------------------------------------------------------------ slab decomposition: ghost exchange
// Fixed-capacity buffers with the count in a header row, so that no count travels to the host.  blockIdx.y = side
// (0: towards the left neighbour, 1: towards the right one).
constexpr int kHaloCols = SPH_HALO_COLS;

__global__ void __launch_bounds__(kBlock)
halo_pack_kernel(sph_fields f, const int32_t *__restrict__ idx_left, const int32_t *__restrict__ idx_right,
                 uint32_t cap, double *__restrict__ send_left, double *__restrict__ send_right,
                 sph_status *__restrict__ status)
{
    const int side = blockIdx.y;
    const uint32_t have = status->halo_count[side], cnt = have < cap ? have : cap;
    const int32_t *idx = side ? idx_right : idx_left;
    double *rows = side ? send_right : send_left;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        rows[0] = (double)cnt;
        rows[1] = (double)have;
#pragma unroll
        for (int c = 2; c < kHaloCols; ++c) rows[c] = 0.0;
        if (have > cap) atomicOr(&status->flags, SPH_F_HALO_OVERFLOW);
    }
    if (k >= cnt) return;
    const size_t i = (size_t)idx[k];
    double *o = rows + kHaloCols * ((size_t)k + 1);
    o[0] = f.r[3 * i]; o[1] = f.r[3 * i + 1]; o[2] = f.r[3 * i + 2];
    o[3] = f.v[3 * i]; o[4] = f.v[3 * i + 1]; o[5] = f.v[3 * i + 2];
    o[6] = f.m[i]; o[7] = f.h[i]; o[8] = f.t[i];
    o[9] = (double)f.gid[i];
}

// The left neighbour's particles fill the slots first .. first + cL - 1, the right neighbour's the cR behind them.
__global__ void __launch_bounds__(kBlock)
halo_unpack_kernel(sph_fields f, const double *__restrict__ recv_left, const double *__restrict__ recv_right,
                   uint32_t cap, int first, int32_t *__restrict__ n_valid, sph_status *__restrict__ status)
{
    const int side = blockIdx.y;
    uint32_t cL = (uint32_t)recv_left[0], cR = (uint32_t)recv_right[0];
    cL = cL < cap ? cL : cap;
    cR = cR < cap ? cR : cap;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0 && side == 0) {
        *n_valid = first + (int)(cL + cR);
        status->ghost_count[0] = cL;
        status->ghost_count[1] = cR;
    }
    if (k >= (side ? cR : cL)) return;
    const double *o = (side ? recv_right : recv_left) + kHaloCols * ((size_t)k + 1);
    const size_t i = (size_t)first + (side ? cL : 0u) + k;
    f.r[3 * i] = o[0]; f.r[3 * i + 1] = o[1]; f.r[3 * i + 2] = o[2];
    f.v[3 * i] = o[3]; f.v[3 * i + 1] = o[4]; f.v[3 * i + 2] = o[5];
    f.m[i] = o[6]; f.h[i] = o[7]; f.t[i] = o[8];
    f.gid[i] = (int64_t)o[9];
}

__global__ void __launch_bounds__(kBlock)
halo_pack2_kernel(const int32_t *__restrict__ idx_left, const int32_t *__restrict__ idx_right, uint32_t cap,
                  const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ send_left,
                  double *__restrict__ send_right, const sph_status *__restrict__ status)
{
    const int side = blockIdx.y;
    const uint32_t have = status->halo_count[side], cnt = have < cap ? have : cap;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    const size_t i = (size_t)(side ? idx_right : idx_left)[k];
    double *o = (side ? send_right : send_left) + 2 * (size_t)k;
    o[0] = a[i];
    o[1] = b[i];
}

__global__ void __launch_bounds__(kBlock)
halo_unpack2_kernel(const double *__restrict__ recv_left, const double *__restrict__ recv_right, int first,
                    double *__restrict__ a, double *__restrict__ b, const sph_status *__restrict__ status)
{
    const int side = blockIdx.y;
    const uint32_t cL = status->ghost_count[0], cR = status->ghost_count[1];
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (side ? cR : cL)) return;
    const double *o = (side ? recv_right : recv_left) + 2 * (size_t)k;
    const size_t i = (size_t)first + (side ? cL : 0u) + k;
    a[i] = o[0];
    b[i] = o[1];
}

inline int blocks_for(int64_t n, int per) { return (int)((n + per - 1) / per); }

inline int launch_status()
{
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPH_OK : (int)e;
}

// Lanes cooperating on one particle in the density / force passes.  Measured on B200 (256^3):
// 1 lane per particle is fastest (density 2.10 ms; 2 lanes 2.52, 4 lanes 3.69, 8 lanes 5.23):
// splitting a row over lanes makes the index loads touch LPP lines per request and buys nothing
// on the gathers, which cost one L1 wavefront per 32-byte sector either way.  SPH_LPP=1|2|4|8 in
// the environment selects the other instantiations for experiments.
int lanes_per_particle()
{
    static int lpp = 0;
    if (!lpp) {
        const char *e = getenv("SPH_LPP");
        lpp = e ? atoi(e) : 1;
        if (lpp != 1 && lpp != 2 && lpp != 4 && lpp != 8) lpp = 1;
    }
    return lpp;
}

// per device: one process may drive several GPUs
int sm_count()
{
    static int sm_of[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev >= 0 && dev < 64 ? dev : 0;
    if (!sm_of[slot]) {
        cudaDeviceGetAttribute(&sm_of[slot], cudaDevAttrMultiProcessorCount, dev);
        if (sm_of[slot] <= 0) sm_of[slot] = 148;
    }
    return sm_of[slot];
}

constexpr double kCellTarget = 8.0;    // particles per cell the planner widens sparse grids towards (0: off)

}  // namespace

// ====================================================================== C ABI
extern "C" {

#define SPH_STR2(x) #x
#define SPH_STR(x) SPH_STR2(x)
const char *sph_version(void) { return "pyticles_b200 0.3 (sm_100a, abi " SPH_STR(SPH_ABI_VERSION) ")"; }

int64_t sph_scan_tmp_elems(uint32_t ncode)
{
    return ((int64_t)ncode + kScanTile - 1) / kScanTile + 3;
}

int64_t sph_nbr_elems(int32_t n, int32_t max_nbrs)
{
    return (((int64_t)n + 31) / 32) * 32 * (int64_t)max_nbrs;
}

int64_t sph_group_tab_elems(const sph_grid *g)
{
    return g ? sph_tiles::group_tab_elems(g) : 0;
}

int sph_group_table(const sph_grid *g, const sph_buffers *b, void *stream)
{
    if (!g || !b) return SPH_E_BADARG;
    if (!b->group_tab || sph_tiles::group_tab_elems(g) == 0) return SPH_OK;
    return sph_tiles::fill_group_table(g, const_cast<uint32_t *>(b->group_tab), (cudaStream_t)stream);
}

static uint32_t host_pdep(uint32_t v, uint32_t mask)
{
    uint32_t r = 0;
    for (uint32_t bit = 1; mask; bit <<= 1) {
        const uint32_t low = mask & (0u - mask);
        if (v & 1u) r |= low;
        v >>= 1;
        mask ^= low;
        (void)bit;
    }
    return r;
}

static int grid_set_codes(sph_grid *g)
{
    int total_bits = 0;
    unsigned long long ncode = 1;
    for (int d = 0; d < 3; ++d) {
        int b = 0;
        while ((1 << b) < g->ncl[d]) ++b;
        g->lb[d] = (uint32_t)(b < 3 ? b : 3);                 // blocks of at most 8 layers per dimension
        g->nblk[d] = (uint32_t)((g->ncl[d] + (1 << g->lb[d]) - 1) >> g->lb[d]);
        total_bits += (int)g->lb[d];
        ncode *= g->nblk[d];
    }
    // generalised Morton code inside a block: deal the bits of x, y, z round-robin while each has bits left
    g->mask[0] = g->mask[1] = g->mask[2] = 0;
    uint32_t used[3] = {0, 0, 0};
    int out = 0;
    while (out < total_bits)
        for (int d = 0; d < 3; ++d)
            if (used[d] < g->lb[d]) { g->mask[d] |= 1u << out; ++out; ++used[d]; }
    g->lbits = (uint32_t)total_bits;
    ncode <<= total_bits;
    if (ncode >= (1ull << 31)) return SPH_E_TOOBIG;
    g->ncode = (uint32_t)ncode;
    // local code bits of the last layer (inside the last block) per dimension
    for (int d = 0; d < 3; ++d)
        g->top[d] = host_pdep((uint32_t)((g->ncl[d] - 1) & ((1 << g->lb[d]) - 1)), g->mask[d]);
    g->magic0 = (uint32_t)(((1ull << 32) + g->nblk[0] - 1) / g->nblk[0]);     // unused when nblk == 1
    g->magic1 = (uint32_t)(((1ull << 32) + g->nblk[1] - 1) / g->nblk[1]);
    return SPH_OK;
}

int sph_grid_plan(const double box[3], double cutoff, double tolerance, int64_t n_hint,
                  const double *occ_lo, const double *occ_hi, sph_grid *g)
{
    if (!box || !g) return SPH_E_BADARG;
    const double thr = cutoff * cutoff + tolerance * tolerance;    // neighbour_list.py:155-157
    if (!(thr > 0.0) || !isfinite(thr)) return SPH_E_GRID;
    const double rl = sqrt(thr);
    const double wmin = rl * (1.0 + 1.0 / 1048576.0);
    int ncmax[3];
    for (int d = 0; d < 3; ++d) {
        if (!(box[d] > 0.0) || !isfinite(box[d])) return SPH_E_GRID;
        g->box[d] = box[d];
        double q = floor(box[d] / wmin);
        if (q < 1.0) q = 1.0;
        if (q > 4096.0) q = 4096.0;
        ncmax[d] = (int)q;
        g->nc[d] = ncmax[d];
    }
    g->thr = thr;
    // keep the cell table O(n): coarsen the dimensions whose cells are mostly empty
    if (n_hint > 0) {
        const double budget = fmax(32768.0, 2.0 * (double)n_hint);
        for (int iter = 0; iter < 64; ++iter) {
            const double cells = (double)g->nc[0] * g->nc[1] * g->nc[2];
            if (cells <= budget) break;
            int best = -1;
            double best_waste = 1.0;
            for (int d = 0; d < 3; ++d) {
                if (g->nc[d] <= 3) continue;
                double occ = g->nc[d];
                if (occ_lo && occ_hi) {
                    const double w = box[d] / g->nc[d];
                    occ = floor((occ_hi[d] - occ_lo[d]) / w) + 2.0;
                    if (!(occ >= 1.0)) occ = 1.0;
                    if (occ > g->nc[d]) occ = g->nc[d];
                }
                const double waste = g->nc[d] / occ;
                if (waste > best_waste * 1.5) { best_waste = waste; best = d; }
            }
            if (best < 0) break;
            g->nc[best] = g->nc[best] / 2 < 3 ? 3 : g->nc[best] / 2;
        }
    }
    // Few particles per cell: the neighbour pass pays its per-cell work (window staging, scans, row write-out) for
    // too few particles -- a 256^3 box at cutoff 1.5 (3.4 per cell) builds slower than at cutoff 2 (8 per cell) although
    // it tests fewer candidates (profiles/r1e_nl_sweep.txt).  Widen the cells of the occupied dimensions towards
    // kCellTarget particles per cell; any width >= the list radius keeps the cells a superset filter.
    // SPH_CELL_TARGET in the environment overrides the target (0: always the narrowest cells).
    if (n_hint > 0) {
        static double target = -1.0;
        if (target < 0.0) {
            const char *e = getenv("SPH_CELL_TARGET");
            target = e ? atof(e) : kCellTarget;
            if (!(target >= 0.0) || target > 64.0) target = kCellTarget;
        }
        double occ[3], cells = 1.0;
        int wide = 0;
        for (int d = 0; d < 3; ++d) {
            occ[d] = g->nc[d];
            if (occ_lo && occ_hi) {
                occ[d] = floor((occ_hi[d] - occ_lo[d]) / (box[d] / g->nc[d])) + 1.0;
                if (!(occ[d] >= 1.0)) occ[d] = 1.0;
                if (occ[d] > g->nc[d]) occ[d] = g->nc[d];
            }
            cells *= occ[d];
            if (occ[d] >= 4.0) ++wide;
        }
        const double mean = (double)n_hint / cells;
        if (target > 0.0 && wide > 0 && mean < 0.625 * target) {
            const double f = pow(mean / target, 1.0 / wide);
            for (int d = 0; d < 3; ++d) {
                if (occ[d] < 4.0) continue;
                int nc = (int)floor(g->nc[d] * f);
                g->nc[d] = nc < 3 ? 3 : nc;
            }
        }
    }
    double wmax = 0.0;
    for (int d = 0; d < 3; ++d) {
        g->w[d] = box[d] / g->nc[d];
        g->inv_w[d] = g->nc[d] / box[d];
        g->lo[d] = 0;
        g->ncl[d] = g->nc[d];
        g->wrap[d] = 1;
        if (g->w[d] > wmax) wmax = g->w[d];
    }
    if (grid_set_codes(g) != SPH_OK) return SPH_E_TOOBIG;
    // fp32 pre-filter band.  Cell-relative coordinates carry an absolute error of at most
    // ~4 * 2^-24 * wmax per component after the shift and the subtraction; rsq inherits
    // 2*sqrt(3)*rl*err + 3*err^2 plus ~8*2^-24 relative from its own arithmetic.  Use 4x that.
    const double u = 1.0 / 16777216.0;
    const double err = 8.0 * u * wmax;
    const double band = 4.0 * (2.0 * 1.7320508 * (rl + err) * err + 3.0 * err * err + 8.0 * u * thr);
    float tin = (float)(thr - band), tout = (float)(thr + band);
    tin = nextafterf(tin, -INFINITY);
    tout = nextafterf(tout, INFINITY);
    if (!(tin > 0.0f)) tin = 0.0f;
    g->thr_in = tin;
    g->thr_out = tout;
    return SPH_OK;
}

int sph_grid_restrict_x(sph_grid *g, int32_t first_layer, int32_t n_layers)
{
    if (!g || n_layers <= 0 || n_layers > g->nc[0]) return SPH_E_BADARG;
    first_layer %= g->nc[0];
    if (first_layer < 0) first_layer += g->nc[0];
    g->lo[0] = first_layer;
    g->ncl[0] = n_layers;
    g->wrap[0] = (n_layers == g->nc[0]) ? 1 : 0;
    return grid_set_codes(g);
}

int sph_status_reset(sph_status *d_status, void *stream)
{
    if (!d_status) return SPH_E_BADARG;
    status_reset_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_status);
    return launch_status();
}

int sph_exclusive_scan_u32(const uint32_t *d_in, uint32_t *d_out, uint32_t *d_tmp, int64_t n,
                           void *stream)
{
    if (!d_in || !d_out || !d_tmp || n <= 0) return SPH_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ntiles = (n + kScanTile - 1) / kScanTile;
    scan_tile_sums<<<(unsigned)ntiles, kBlock, 0, s>>>(d_in, n, d_tmp);
    scan_tile_offsets<<<1, kBlock, 0, s>>>(d_tmp, ntiles);
    scan_tile_apply<<<(unsigned)ntiles, kBlock, 0, s>>>(d_in, d_out, n, d_tmp, ntiles);
    return launch_status();
}

static bool cells_args_ok(const sph_grid *g, const sph_buffers *b, const double *d_r)
{
    return g && b && d_r && b->n >= 0 && b->cell_count && b->cell_start && b->scan_tmp && b->code && b->rank &&
           b->perm && b->status;
}

int sph_cells_begin(const sph_grid *g, const sph_buffers *b, const double *d_r, int32_t first, int32_t count,
                    int32_t *d_idx_left, int32_t *d_idx_right, int32_t cap, void *stream)
{
    if (!cells_args_ok(g, b, d_r) || first < 0 || count < 0 || first + count > b->n || cap < 0) return SPH_E_BADARG;
    if ((d_idx_left == nullptr) != (d_idx_right == nullptr)) return SPH_E_BADARG;
    if (d_idx_left && (g->wrap[0] || g->ncl[0] < 4)) return SPH_E_BADARG;     // needs a restricted slab grid
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(b->cell_count, 0, sizeof(uint32_t) * ((size_t)g->ncode + 1), s);
    const HaloLists hl = {d_idx_left, d_idx_right, (uint32_t)cap};
    if (count > 0)
        // no n_valid here: it is written by sph_halo_unpack AFTER this pass (the range binned first holds particles)
        bin_kernel<<<blocks_for(count, kBlock), kBlock, 0, s>>>(*g, d_r, first, count, nullptr, b->n_owned,
                                                                b->cell_count, b->code, b->rank, b->status, hl);
    return launch_status();
}

int sph_cells_add(const sph_grid *g, const sph_buffers *b, const double *d_r, int32_t first, int32_t count,
                  void *stream)
{
    if (!cells_args_ok(g, b, d_r) || first < 0 || count < 0 || first + count > b->n) return SPH_E_BADARG;
    const HaloLists hl = {nullptr, nullptr, 0u};
    if (count > 0)
        bin_kernel<<<blocks_for(count, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
            *g, d_r, first, count, b->n_valid, b->n_owned, b->cell_count, b->code, b->rank, b->status, hl);
    return launch_status();
}

int sph_cells_finish(const sph_grid *g, const sph_buffers *b, void *stream)
{
    if (!g || !b || !b->cell_count || !b->cell_start || !b->scan_tmp || !b->code || !b->rank || !b->perm)
        return SPH_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    // one cell more than the grid has: the spare cell of the unused ghost slots
    int rc = sph_exclusive_scan_u32(b->cell_count, b->cell_start, b->scan_tmp, (int64_t)g->ncode + 1, stream);
    if (rc != SPH_OK) return rc;
    if (b->n > 0) {
        scatter_kernel<<<blocks_for(b->n, kBlock), kBlock, 0, s>>>(b->n, b->code, b->rank, b->cell_start, b->perm);
        cell_sort_kernel<<<blocks_for(g->ncode, kBlock), kBlock, 0, s>>>(g->ncode, b->cell_start, b->perm, b->n_owned,
                                                                        b->sort_key);
    }
    return launch_status();
}

int sph_cells_build(const sph_grid *g, const sph_buffers *b, const double *d_r, void *stream)
{
    if (!cells_args_ok(g, b, d_r)) return SPH_E_BADARG;
    const int rc = sph_cells_begin(g, b, d_r, 0, b->n, nullptr, nullptr, 0, stream);
    return rc != SPH_OK ? rc : sph_cells_finish(g, b, stream);
}

int sph_gather(const sph_grid *g, const sph_buffers *b, const double *d_r, const double *d_v,
               const double *d_m, void *stream)
{
    if (!g || !b || !d_r || !d_v || !d_m || !b->pos4 || !b->vel4 || !b->rel4 || !b->perm || !b->cnt || !b->cell_start)
        return SPH_E_BADARG;
    if (b->n > 0)
        gather_kernel<<<blocks_for(b->n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
            *g, b->n, b->perm, d_r, d_v, d_m, b->pos4, b->vel4, b->rel4, b->cnt, b->cell_start);
    return launch_status();
}

static int nlist_general(const sph_grid *g, const sph_buffers *b, int only_fallback, cudaStream_t s)
{
    static bool configured[64] = {};                     // the attribute belongs to the function ON a device
    const size_t smem = sizeof(float4) * kNlWin * kNlWarps;
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev >= 0 && dev < 64 ? dev : 0;
    if (!configured[slot]) {
        cudaFuncSetAttribute(nlist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(nlist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[slot] = true;
    }
    const bool small = g->ncl[0] < 3 || g->ncl[1] < 3 || g->ncl[2] < 3;
    const int64_t warps_needed = g->ncode;
    int64_t blocks = (warps_needed + kNlWarps - 1) / kNlWarps;
    const int64_t cap = (int64_t)sm_count() * 3 * 8;       // a few resident waves, grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (small)
        nlist_kernel<true><<<(unsigned)blocks, kNlWarps * 32, smem, s>>>(
            *g, b->n, b->max_nbrs, b->cell_start, b->rel4, b->pos4, b->nbr, b->cnt, b->status, only_fallback, b->perm,
            b->n_owned);
    else
        nlist_kernel<false><<<(unsigned)blocks, kNlWarps * 32, smem, s>>>(
            *g, b->n, b->max_nbrs, b->cell_start, b->rel4, b->pos4, b->nbr, b->cnt, b->status, only_fallback, b->perm,
            b->n_owned);
    return launch_status();
}

int sph_nlist_build(const sph_grid *g, const sph_buffers *b, void *stream)
{
    if (!g || !b || !b->nbr || !b->cnt || !b->rel4 || !b->pos4 || !b->cell_start || !b->status) return SPH_E_BADARG;
    if (b->max_nbrs <= 0) return SPH_E_BADARG;
    if (b->n == 0) return SPH_OK;
    cudaStream_t s = (cudaStream_t)stream;
    // the cell-group kernel first; the general kernel behind it returns at once unless that one gave up.
    // SPH_TILES in the environment: 0 general kernel only, 1 (default) scalar cell-group kernel, 2 its tensor-core
    // pre-filter variant (tests, A/B timing)
    const char *env = getenv("SPH_TILES");                 // read at every call: tests switch it
    const int mode = env ? atoi(env) : 1;
    const bool mma = mode == 2 && sph_tiles_mma::eligible(g, b);
    const bool tiles = mma || (mode != 0 && sph_tiles::eligible(g, b));
    if (tiles) {
        flags_clear_kernel<<<1, 1, 0, s>>>(b->status, SPH_F_TILE_FALLBACK);
        const int rc = mma ? sph_tiles_mma::launch_list(g, b, s) : sph_tiles::launch_list(g, b, s);
        if (rc != SPH_OK) return rc;
    }
    return nlist_general(g, b, tiles ? 1 : 0, s);
}

int sph_density_eos(const sph_grid *g, const sph_buffers *b, const sph_eos *eos,
                    const double *d_h_orig, int h_uniform, int list_fresh, int use_hlr,
                    double *d_rho, double *d_p, double *d_pco, double *d_u, double *d_t, void *stream)
{
    if (!g || !b || !eos || !d_h_orig || !d_rho) return SPH_E_BADARG;
    if (use_hlr < 0 || use_hlr > 2 || (use_hlr != 1 && (!d_p || !d_pco || !d_u || !d_t))) return SPH_E_BADARG;
    if (b->n == 0) return SPH_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int lpp = lanes_per_particle();
#define SPH_LAUNCH_DENSITY(U, L)                                                                              \
    density_kernel<U, L><<<blocks_for((int64_t)b->n * L, kPPBlock), kPPBlock, 0, s>>>(                        \
        *g, b->n, b->max_nbrs, b->pos4, b->vel4, b->rel4, b->perm, b->nbr, b->cnt, b->status, d_h_orig, *eos, \
        list_fresh, use_hlr, b->n_owned, d_rho, d_p, d_pco, d_u, d_t)
    if (h_uniform) {
        if (lpp == 1) SPH_LAUNCH_DENSITY(true, 1);
        else if (lpp == 2) SPH_LAUNCH_DENSITY(true, 2);
        else if (lpp == 8) SPH_LAUNCH_DENSITY(true, 8);
        else SPH_LAUNCH_DENSITY(true, 4);
    } else {
        SPH_LAUNCH_DENSITY(false, 1);
    }
#undef SPH_LAUNCH_DENSITY
    return launch_status();
}

int sph_force(const sph_grid *g, const sph_buffers *b, const double *d_press, const double *d_rho,
              const double *d_h_orig, int h_uniform, int list_fresh, double fcutoff, int dim, int first_force,
              int part, double *d_vdot, double *d_udot, void *stream)
{
    if (!g || !b || !d_h_orig || !d_vdot || !d_udot || part < 0 || part > 2) return SPH_E_BADARG;
    if ((d_press == nullptr) != (d_rho == nullptr)) return SPH_E_BADARG;
    if (b->n == 0) return SPH_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_for(b->n, kBlock);
    if (d_press)
        pressure_term_kernel<<<nb, kBlock, 0, s>>>(b->n, 0, b->perm, d_press, d_rho, b->vel4);
    const double fcutsq = fcutoff * fcutoff;                // forces.py:36
    const int lpp = lanes_per_particle();

#define SPH_LAUNCH_FORCE(U, L)                                                                                 \
    force_kernel<U, L><<<blocks_for((int64_t)b->n * L, kPPBlock), kPPBlock, 0, s>>>(                           \
        *g, b->n, b->max_nbrs, b->pos4, b->vel4, b->rel4, b->perm, b->nbr, b->cnt, b->status, d_h_orig,        \
        list_fresh, fcutsq, dim, b->n_owned, first_force ? 1 : 0, part, d_vdot, d_udot)
    if (part == 2 && b->n_owned > 0 && !g->wrap[0] && g->ncl[0] >= 3) {
        // a slab's boundary layers through the cell table: 16 lanes per cell of the two layers
        const int64_t lanes = (int64_t)(g->ncl[0] - 2 > 1 ? 2 : 1) * g->ncl[1] * g->ncl[2] * 16;
#define SPH_LAUNCH_EDGE(U)                                                                                      \
    force_edge_kernel<U><<<blocks_for(lanes, kPPBlock), kPPBlock, 0, s>>>(                                      \
        *g, b->max_nbrs, b->pos4, b->vel4, b->rel4, b->perm, b->nbr, b->cnt, b->cell_start, b->status, d_h_orig, \
        list_fresh, fcutsq, dim, b->n_owned, first_force ? 1 : 0, d_vdot, d_udot)
        if (h_uniform) SPH_LAUNCH_EDGE(true);
        else SPH_LAUNCH_EDGE(false);
#undef SPH_LAUNCH_EDGE
        return launch_status();
    }
    if (h_uniform) {
        if (lpp == 1) SPH_LAUNCH_FORCE(true, 1);
        else if (lpp == 2) SPH_LAUNCH_FORCE(true, 2);
        else if (lpp == 8) SPH_LAUNCH_FORCE(true, 8);
        else SPH_LAUNCH_FORCE(true, 4);
    } else {
        SPH_LAUNCH_FORCE(false, 1);
    }
#undef SPH_LAUNCH_FORCE
    return launch_status();
}

int sph_pressure_term(const sph_buffers *b, const double *d_press, const double *d_rho, int32_t first_orig,
                      void *stream)
{
    if (!b || !d_press || !d_rho || !b->perm || !b->vel4 || first_orig < 0) return SPH_E_BADARG;
    if (b->n > 0)
        pressure_term_kernel<<<blocks_for(b->n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
            b->n, first_orig, b->perm, d_press, d_rho, b->vel4);
    return launch_status();
}

int sph_conduction(const sph_grid *g, const sph_buffers *b, const double *d_jq, const double *d_rho,
                   const double *d_h_orig, int h_uniform, int list_fresh, double *d_aux4, double *d_udot,
                   void *stream)
{
    if (!g || !b || !d_jq || !d_rho || !d_h_orig || !d_aux4 || !d_udot) return SPH_E_BADARG;
    if (b->n == 0) return SPH_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_for(b->n, kBlock);
    flux_term_kernel<<<nb, kBlock, 0, s>>>(b->n, b->perm, d_jq, d_rho, d_aux4);
    if (h_uniform)
        conduction_kernel<true><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, d_aux4, b->rel4, b->perm, b->nbr,
                                                      b->cnt, b->status, d_h_orig, list_fresh, d_udot);
    else
        conduction_kernel<false><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, d_aux4, b->rel4, b->perm, b->nbr,
                                                       b->cnt, b->status, d_h_orig, list_fresh, d_udot);
    return launch_status();
}

int sph_pairs_count(const sph_buffers *b, uint32_t *d_row_count, void *stream)
{
    if (!b || !d_row_count) return SPH_E_BADARG;
    if (b->n > 0)
        pairs_count_kernel<<<blocks_for(b->n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
            b->n, b->max_nbrs, b->perm, b->nbr, b->cnt, d_row_count);
    return launch_status();
}

int sph_pairs_fill(const sph_buffers *b, const int64_t *d_row_start, int32_t *d_iap, int64_t cap_pairs, void *stream)
{
    if (!b || !d_row_start || (!d_iap && cap_pairs > 0)) return SPH_E_BADARG;
    if (b->n > 0 && cap_pairs > 0)
        pairs_fill_kernel<<<blocks_for(b->n, 128), 128, 0, (cudaStream_t)stream>>>(
            b->n, b->max_nbrs, b->perm, b->nbr, b->cnt, d_row_start, d_iap, cap_pairs);
    return launch_status();
}

int sph_separations(const double box[3], const int32_t *d_iap, int64_t nip, const double *d_r,
                    const double *d_v, double *d_drij, double *d_rij, double *d_rsq, double *d_dv, void *stream)
{
    if (!box || nip < 0) return SPH_E_BADARG;
    if (nip == 0) return SPH_OK;
    if (!d_iap || !d_r || !d_v || !d_drij || !d_rij || !d_rsq || !d_dv) return SPH_E_BADARG;
    separations_kernel<<<blocks_for(nip, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        box[0], box[1], box[2], d_iap, nip, d_r, d_v, d_drij, d_rij, d_rsq, d_dv);
    return launch_status();
}

int sph_pair_kernels(const int32_t *d_iap, int64_t nip, const double *d_rij, const double *d_drij,
                     const double *d_h, double *d_wij, double *d_dwij, void *stream)
{
    if (nip < 0) return SPH_E_BADARG;
    if (nip == 0) return SPH_OK;
    if (!d_iap || !d_rij || !d_drij || !d_h || !d_wij || !d_dwij) return SPH_E_BADARG;
    pair_kernels_kernel<<<blocks_for(nip, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        d_iap, nip, d_rij, d_drij, d_h, d_wij, d_dwij);
    return launch_status();
}

int sph_compress(const sph_grid *g, const sph_buffers *b, void *stream)
{
    if (!g || !b || !b->nbr || !b->cnt || !b->pos4) return SPH_E_BADARG;
    if (b->n > 0)
        compress_kernel<<<blocks_for(b->n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
            *g, b->n, b->max_nbrs, b->pos4, b->nbr, b->cnt);
    return launch_status();
}

int sph_ponder_rebuild(const double *d_r_old, const double *d_r, int32_t n, double tol_sq,
                       sph_status *d_status, void *stream)
{
    if (!d_r_old || !d_r || !d_status || n < 0) return SPH_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(&d_status->dsq_max_bits, 0, sizeof(unsigned long long), s);
    if (n > 0) ponder_kernel<<<blocks_for(n, kBlock), kBlock, 0, s>>>(d_r_old, d_r, n, d_status);
    ponder_decide_kernel<<<1, 1, 0, s>>>(d_status, tol_sq);
    return launch_status();
}

static bool fields_ok(const sph_fields *f)
{
    return f && f->r && f->v && f->m && f->h && f->t && f->gid;
}

int sph_halo_pack(const sph_fields *f, const int32_t *d_idx_left, const int32_t *d_idx_right, int32_t cap,
                  double *d_send_left, double *d_send_right, sph_status *d_status, void *stream)
{
    if (!fields_ok(f) || !d_idx_left || !d_idx_right || cap <= 0 || !d_send_left || !d_send_right || !d_status)
        return SPH_E_BADARG;
    const dim3 grid((unsigned)blocks_for(cap, kBlock), 2u);
    halo_pack_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(*f, d_idx_left, d_idx_right, (uint32_t)cap, d_send_left,
                                                                d_send_right, d_status);
    return launch_status();
}

int sph_halo_unpack(const sph_fields *f, const double *d_recv_left, const double *d_recv_right, int32_t cap,
                    int32_t first, int32_t *d_n_valid, sph_status *d_status, void *stream)
{
    if (!fields_ok(f) || !d_recv_left || !d_recv_right || cap <= 0 || first < 0 || !d_n_valid || !d_status)
        return SPH_E_BADARG;
    const dim3 grid((unsigned)blocks_for(cap, kBlock), 2u);
    halo_unpack_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(*f, d_recv_left, d_recv_right, (uint32_t)cap, first,
                                                                  d_n_valid, d_status);
    return launch_status();
}

int sph_halo_pack2(const int32_t *d_idx_left, const int32_t *d_idx_right, int32_t cap, const double *d_a,
                   const double *d_b, double *d_send_left, double *d_send_right, const sph_status *d_status,
                   void *stream)
{
    if (!d_idx_left || !d_idx_right || cap <= 0 || !d_a || !d_b || !d_send_left || !d_send_right || !d_status)
        return SPH_E_BADARG;
    const dim3 grid((unsigned)blocks_for(cap, kBlock), 2u);
    halo_pack2_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(d_idx_left, d_idx_right, (uint32_t)cap, d_a, d_b,
                                                                 d_send_left, d_send_right, d_status);
    return launch_status();
}

int sph_halo_unpack2(const double *d_recv_left, const double *d_recv_right, int32_t cap, int32_t first,
                     double *d_a, double *d_b, const sph_status *d_status, void *stream)
{
    if (!d_recv_left || !d_recv_right || cap <= 0 || first < 0 || !d_a || !d_b || !d_status) return SPH_E_BADARG;
    const dim3 grid((unsigned)blocks_for(cap, kBlock), 2u);
    halo_unpack2_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(d_recv_left, d_recv_right, first, d_a, d_b, d_status);
    return launch_status();
}

int sph_axpy(double *d_x, const double *d_a, const double *d_b, double sc, int64_t len, void *stream)
{
    if (len < 0 || (len > 0 && (!d_x || !d_a || !d_b))) return SPH_E_BADARG;
    if (len > 0) axpy_kernel<<<blocks_for(len, kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_x, d_a, d_b, sc, len);
    return launch_status();
}

int sph_box_apply(const double box[3], int kind, double *d_r, double *d_v, int32_t n, void *stream)
{
    if (!box || n < 0 || (n > 0 && (!d_r || !d_v))) return SPH_E_BADARG;
    if (n > 0) box_kernel<<<blocks_for(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(box[0], box[1], box[2], kind, d_r, d_v, n);
    return launch_status();
}

}  // extern "C"
