// Micro-benchmark (development aid): how fast can the GPU fetch scattered 40-byte rows from pinned HOST memory?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/pcie_gather_bench.cu -o /tmp/pcie_gather && /tmp/pcie_gather
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int C = 10;

__global__ void v1_elem(const float *__restrict__ src, const int *__restrict__ pix, int n, float *__restrict__ dst) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * C; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), c = (int)(i - (long long)r * C);
        dst[i] = src[(long long)pix[r] * C + c];
    }
}
__global__ void v2_row8(const float *__restrict__ src, const int *__restrict__ pix, int n, float *__restrict__ dst) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const float2 *s = reinterpret_cast<const float2 *>(src + (long long)pix[r] * C);
        float2 v[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = s[k];
        float2 *d = reinterpret_cast<float2 *>(dst + (long long)r * C);
#pragma unroll
        for (int k = 0; k < 5; ++k) d[k] = v[k];
    }
}
// 8 lanes per row: the lanes fetch the whole 128-byte line(s) that hold the row with 16-byte loads
__global__ void v3_line(const float *__restrict__ src, const int *__restrict__ pix, int n, float *__restrict__ dst) {
    const int sub = threadIdx.x & 7;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < n; r += (gridDim.x * blockDim.x) >> 3) {
        const long long b0 = (long long)pix[r] * C * 4, b1 = b0 + C * 4;  // byte range of the row
        const long long line0 = b0 & ~127LL;
        for (long long line = line0; line < b1; line += 128) {
            const long long a = line + sub * 16;
            const float4 v = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(src) + a);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long e = (a + 4 * j - b0) / 4;
                if (a + 4 * j >= b0 && e < C) dst[(long long)r * C + e] = vv[j];
            }
        }
    }
}
// 4 lanes per row, 16-byte loads covering only the sectors the row touches (row start rounded down to 16 B)
__global__ void v4_vec16(const float *__restrict__ src, const int *__restrict__ pix, int n, float *__restrict__ dst) {
    const int sub = threadIdx.x & 3;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; r < n; r += (gridDim.x * blockDim.x) >> 2) {
        const long long b0 = (long long)pix[r] * C * 4, b1 = b0 + C * 4;
        const long long a = (b0 & ~15LL) + sub * 16;
        if (a < b1) {
            const float4 v = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(src) + a);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long e = (a + 4 * j - b0) / 4;
                if (a + 4 * j >= b0 && e < C) dst[(long long)r * C + e] = vv[j];
            }
        }
    }
}

int main(int argc, char **argv) {
    const long long P = 19961856;
    const int n = 250000;
    float *h = nullptr;
    CK(cudaHostAlloc(&h, P * C * sizeof(float), cudaHostAllocDefault));
    for (long long i = 0; i < P * C; i += 1024) h[i] = (float)i;
    std::vector<int> pix(n);
    srand(1);
    const int order = argc > 1 ? atoi(argv[1]) : 0;  // 0: sorted by pixel, 1: random, 2: Z-order of 32x32-pixel cells
    const int W = 5472;
    for (int v = 0; v < 10; ++v) {  // 10 views x 25k winners
        std::vector<int> p(n / 10);
        for (auto &x : p) x = (int)(((long long)rand() * 32768 + rand()) % P);
        if (order == 0) std::sort(p.begin(), p.end());
        if (order == 2) {
            auto key = [&](int q) {
                unsigned x = (q % W) / 32, y = (q / W) / 32, k = 0;
                for (int b = 0; b < 10; ++b) k |= ((x >> b) & 1u) << (2 * b) | ((y >> b) & 1u) << (2 * b + 1);
                return k;
            };
            std::sort(p.begin(), p.end(), [&](int a, int b) { return key(a) < key(b); });
        }
        std::copy(p.begin(), p.end(), pix.begin() + v * (n / 10));
    }
    printf("order %d (0 sorted by pixel, 1 random, 2 Z-order of 32x32 cells)\n", order);
    int *d_pix; float *d_dst;
    CK(cudaMalloc(&d_pix, n * 4)); CK(cudaMalloc(&d_dst, (size_t)n * C * 4));
    CK(cudaMemcpy(d_pix, pix.data(), n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct V { const char *name; int id; } vs[] = {{"v1 thread/element 4B", 1}, {"v4 4 lanes/row 16B", 4}};
    for (int rep = 0; rep < 2; ++rep)
        for (auto &v : vs) {
            for (int blocks : {148 * 8, 148 * 32}) {
                CK(cudaMemset(d_dst, 0, (size_t)n * C * 4));
                CK(cudaEventRecord(e0));
                for (int it = 0; it < 5; ++it) {
                    if (v.id == 1) v1_elem<<<blocks, 256>>>(h, d_pix, n, d_dst);
                    if (v.id == 2) v2_row8<<<blocks, 256>>>(h, d_pix, n, d_dst);
                    if (v.id == 3) v3_line<<<blocks, 256>>>(h, d_pix, n, d_dst);
                    if (v.id == 4) v4_vec16<<<blocks, 256>>>(h, d_pix, n, d_dst);
                }
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
                if (rep == 1) printf("%-28s blocks %5d: %7.3f ms per 250k rows = %6.1f M rows/s, %5.2f GB/s payload\n", v.name, blocks, ms, n / ms / 1e3, n * 40.0 / ms / 1e6);
            }
        }
    return 0;
}
