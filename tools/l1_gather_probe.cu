// Probe: what bounds scattered 32-byte row gathers through L1 on this GPU -- 32-byte sectors or 128-byte lines?
//   mode 0: every lane loads one random 32-byte row (32 lines per load instruction)            [density pass]
//   mode 1: every lane loads two random 32-byte rows from two arrays (2 x 32 lines)            [force pass]
//   mode 2: lanes 2i, 2i+1 load the two halves of one random 64-byte row (16 lines per load)   [lane-pair idea]
//   mode 3: every lane loads both halves of one random 64-byte row with two loads (2 x 32 lookups of 32 lines)
//   mode 4/5/6: G = 2/4/8 lanes share a 128-byte LINE but read different 32-byte sectors of it (lines 32/G, sectors 32)
//   mode 7/8/9: G = 2/4/8 lanes read the SAME sector (lines and sectors 32/G)
//     -- together they say whether L1 charges a gather by the line or by the sector (added at the end of round 1,
//        not yet run: profiles/r1e_l1_gather_probe.txt holds modes 0-3 only)
// Rows are picked like SPH neighbours: random within a window of W rows around the lane's own row, so the
// working set of a block stays in L1/L2.  Prints rows/s and bytes/s per mode.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/l1_gather_probe.cu -o /tmp/l1probe && /tmp/l1probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void load4(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

__device__ __forceinline__ uint32_t rng(uint32_t &s)
{
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return s;
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(const double *__restrict__ A, const double *__restrict__ B, int n, int W,
                                             int iters, double *__restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t s = 0x9E3779B9u * (uint32_t)(MODE == 2 ? (t >> 1) + 1 : t + 1);
    const int self = MODE == 2 ? (t >> 1) : t;
    double acc = 0.0;
    for (int k = 0; k < iters; ++k) {
        int j = self + (int)(rng(s) % (uint32_t)W) - W / 2;
        j = j < 0 ? j + n : (j >= n ? j - n : j);
        double a, b, c, d;
        if (MODE == 0) {
            load4(A + 4 * (size_t)j, a, b, c, d);
            acc += a + b + c + d;
        } else if (MODE == 1) {
            double e, f, g, h;
            load4(A + 4 * (size_t)j, a, b, c, d);
            load4(B + 4 * (size_t)j, e, f, g, h);
            acc += (a + b + c + d) + (e + f + g + h);
        } else if (MODE == 2) {
            load4(A + 8 * (size_t)j + 4 * (lane & 1), a, b, c, d);      // A holds 64-byte rows here
            acc += a + b + c + d;
        } else if (MODE == 3) {
            double e, f, g, h;
            load4(A + 8 * (size_t)j, a, b, c, d);
            load4(A + 8 * (size_t)j + 4, e, f, g, h);
            acc += (a + b + c + d) + (e + f + g + h);
        } else {
            // groups of G lanes agree on one random line (the group leader's j, rounded to 4 rows); mode 4-6: lane g
            // of the group reads sector g % 4 of it (G = 8: two lanes per sector), mode 7-9: all read sector 0
            constexpr int G = (MODE == 4 || MODE == 7) ? 2 : ((MODE == 5 || MODE == 8) ? 4 : 8);
            const int leader = __shfl_sync(0xffffffffu, j, lane & ~(G - 1));
            const int line = leader & ~3;
            const int sector = MODE <= 6 ? (lane & (G - 1)) & 3 : 0;
            load4(A + 4 * (size_t)(line + sector), a, b, c, d);
            acc += a + b + c + d;
        }
    }
    out[t] = acc;
}

int main()
{
    const int n = 1 << 24, W = 256, iters = 30;
    double *A, *B, *out;
    cudaMalloc(&A, sizeof(double) * 8 * (size_t)n);
    cudaMalloc(&B, sizeof(double) * 4 * (size_t)n);
    cudaMalloc(&out, sizeof(double) * (size_t)n);
    cudaMemset(A, 0, sizeof(double) * 8 * (size_t)n);
    cudaMemset(B, 0, sizeof(double) * 4 * (size_t)n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[10] = {"one 32 B row per lane            ", "two 32 B rows per lane (2 arrays)",
                             "64 B row split over a lane pair  ", "64 B row, two loads per lane     ",
                             "2 lanes per line, own sectors    ", "4 lanes per line, own sectors    ",
                             "8 lanes per line, 2 per sector   ", "2 lanes per sector               ",
                             "4 lanes per sector               ", "8 lanes per sector               "};
    for (int mode = 0; mode < 10; ++mode) {
        const int threads = n;                        // mode 2: n/2 rows-gatherers, two lanes each
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<threads / 256, 256>>>(A, B, n, W, iters, out);
            if (mode == 1) probe<1><<<threads / 256, 256>>>(A, B, n, W, iters, out);
            if (mode == 2) probe<2><<<threads / 256, 256>>>(A, B, n / 2, W, iters, out);
            if (mode == 3) probe<3><<<threads / 256, 256>>>(A, B, n, W, iters, out);
            if (mode == 4) probe<4><<<threads / 256, 256>>>(A, B, n, W, iters, out);
            if (mode == 5) probe<5><<<threads / 256, 256>>>(A, B, n, W, iters, out);
            if (mode == 6) probe<6><<<threads / 256, 256>>>(A, B, n, W, iters, out);
            if (mode == 7) probe<7><<<threads / 256, 256>>>(A, B, n, W, iters, out);
            if (mode == 8) probe<8><<<threads / 256, 256>>>(A, B, n, W, iters, out);
            if (mode == 9) probe<9><<<threads / 256, 256>>>(A, B, n, W, iters, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double rows = (mode == 2 ? 0.5 : 1.0) * (double)n * iters;          // gathered neighbour rows (modes 4-9: loads)
        const double bytes = rows * ((mode == 0 || mode >= 4) ? 32.0 : 64.0);
        printf("mode %d  %s  %.3f ms  %.3e rows/s  %.1f GB/s  %.2f bytes/clk/SM (148 SMs, 1.965 GHz)\n", mode, names[mode],
               best, rows / (best * 1e-3), bytes / (best * 1e-3) / 1e9, bytes / (best * 1e-3) / 148.0 / 1.965e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
