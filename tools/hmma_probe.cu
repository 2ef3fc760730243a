// Probe: issue rate of the legacy warp-level tensor instruction mma.sync.m16n8k16 (f16 inputs, f32 accumulate; SASS HMMA.16816.F32)
// on this GPU, alone and with the per-tile work of a distance pre-filter around it (ldmatrix of the A tile, sign-bit
// funnel shifts and a running |d| minimum on the four accumulators).  Round 2 question: is a tensor-core pre-filter
// of candidate pairs affordable inside the neighbour pass?  Prints tiles/s per SM-clock.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/hmma_probe.cu -o /tmp/hmma && /tmp/hmma
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float &d0, float &d1, float &d2, float &d3, uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d0), "=f"(d1), "=f"(d2), "=f"(d3)
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f), "f"(0.f), "f"(0.f), "f"(0.f));
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, const uint32_t *__restrict__ in, uint32_t *__restrict__ out)
{
    __shared__ __align__(128) uint32_t A[64 * 128];            // 64 tiles of 16 x 16 halves (512 B each)
    for (int i = threadIdx.x; i < 64 * 128; i += blockDim.x) A[i] = in[i & 1023];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t b0 = in[lane], b1 = in[lane + 32];
    uint32_t a0 = in[lane + 64], a1 = in[lane + 96], a2 = in[lane + 128], a3 = in[lane + 160];
    uint32_t m0 = 0, m1 = 0;
    float near = 1e30f, acc = 0.f;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(A) + (lane & 15) * 32 + (lane >> 4) * 16;
    for (int k = 0; k < iters; ++k) {
#pragma unroll 4
        for (int t = 0; t < 16; ++t) {
            float d0, d1, d2, d3;
            if (MODE >= 1) {
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(base + ((t + k) & 63) * 512));
            }
            mma16816(d0, d1, d2, d3, a0, a1, a2, a3, b0, b1);
            if (MODE >= 2) {
                m0 = __funnelshift_l(__float_as_uint(d0), m0, 1);
                m0 = __funnelshift_l(__float_as_uint(d1), m0, 1);
                m1 = __funnelshift_l(__float_as_uint(d2), m1, 1);
                m1 = __funnelshift_l(__float_as_uint(d3), m1, 1);
                near = fminf(near, fminf(fminf(fabsf(d0), fabsf(d1)), fminf(fabsf(d2), fabsf(d3))));
            } else {
                acc += d0 + d1 + d2 + d3;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = m0 ^ m1 ^ __float_as_uint(near) ^ __float_as_uint(acc);
}

int main()
{
    uint32_t *in, *out;
    cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0x3c, 4096 * 4);
    const int blocks = 148 * 8, iters = 2000;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[3] = {"HMMA.16816.F32 only                         ", "ldmatrix.x4 + HMMA                          ",
                            "ldmatrix.x4 + HMMA + 4 sign shifts + |d| min"};
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<blocks, 256>>>(iters, in, out);
            if (mode == 1) probe<1><<<blocks, 256>>>(iters, in, out);
            if (mode == 2) probe<2><<<blocks, 256>>>(iters, in, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double tiles = (double)blocks * 8 * iters * 16;       // warp-level 16x8x16 tiles
        printf("mode %d  %s  %.3f ms  %.3e tiles/s  %.2f SM-clocks per tile (148 SMs, 1.965 GHz)  %.1f dense TFLOP/s\n", mode,
               names[mode], best, tiles / (best * 1e-3), 148.0 * 1.965e9 / (tiles / (best * 1e-3)),
               tiles * 4096.0 / (best * 1e-3) / 1e12);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
