// Micro-benchmark (development aid): scattered reads of ROW_BYTES-byte rows from pinned HOST memory by a GPU kernel,
// for row sizes 1 (uint8 class index), 20 (10 x f16), 32 (one aligned sector), 40 (10 x f32), 64 (aligned line half).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/pcie_rows_bench.cu -o /tmp/pcie_rows && /tmp/pcie_rows
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <cstdint>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// one thread per 4-byte word of a row (1-byte rows: one thread per row)
__global__ void k_rows(const unsigned char *__restrict__ src, const int *__restrict__ pix, int n, int row_bytes,
                       unsigned *__restrict__ dst) {
    const int words = row_bytes >= 4 ? row_bytes / 4 : 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * words;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / words), c = (int)(i - (long long)r * words);
        const long long off = (long long)pix[r] * row_bytes;
        dst[i] = row_bytes >= 4 ? *reinterpret_cast<const unsigned *>(src + off + 4 * c) : (unsigned)src[off];
    }
}

int main(int argc, char **argv) {
    const int n_img = argc > 1 ? atoi(argv[1]) : 1;  // rows are spread over n_img x 19.96 M pixels of pinned memory
    const long long P = 19961856LL * n_img;
    const int n = 250000;
    unsigned char *h = nullptr;
    const int thp = argc > 2 ? atoi(argv[2]) : 0;  // 1: transparent huge pages (mmap + MADV_HUGEPAGE) + cudaHostRegister
    if (thp == 3 || thp == 4) {  // ordinary malloc'ed memory: only readable by the GPU where the driver offers HMM / ATS
        int pma = 0, hpt = 0;
        CK(cudaDeviceGetAttribute(&pma, cudaDevAttrPageableMemoryAccess, 0));
        CK(cudaDeviceGetAttribute(&hpt, cudaDevAttrPageableMemoryAccessUsesHostPageTables, 0));
        printf("pageableMemoryAccess = %d, usesHostPageTables = %d\n", pma, hpt);
        if (!pma) return 0;
        h = (unsigned char *)malloc((size_t)P * 64);
        for (long long i = 0; i < P * 64; i += 4096) h[i] = (unsigned char)i;
        if (thp == 4) {
            cudaError_t e1 = cudaMemAdvise(h, (size_t)P * 64, cudaMemAdviseSetPreferredLocation, cudaCpuDeviceId);
            cudaError_t e2 = cudaMemAdvise(h, (size_t)P * 64, cudaMemAdviseSetAccessedBy, 0);
            printf("malloc + advise: preferred location %s, accessed by %s\n", cudaGetErrorString(e1), cudaGetErrorString(e2));
            (void)cudaGetLastError();
        } else {
            printf("malloc, no advice\n");
        }
    } else if (thp == 2) {  // managed memory that prefers the host and is mapped into the GPU: accessed over PCIe, not migrated
        CK(cudaMallocManaged(&h, (size_t)P * 64));
        CK(cudaMemAdvise(h, (size_t)P * 64, cudaMemAdviseSetPreferredLocation, cudaCpuDeviceId));
        CK(cudaMemAdvise(h, (size_t)P * 64, cudaMemAdviseSetAccessedBy, 0));
        for (long long i = 0; i < P * 64; i += 4096) h[i] = (unsigned char)i;
        printf("managed memory, preferred location = host, accessed by GPU 0\n");
    } else if (thp) {
        const size_t bytes = ((size_t)P * 64 + (2u << 20) - 1) / (2u << 20) * (2u << 20);
        void *m = mmap(nullptr, bytes + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) { printf("mmap failed\n"); return 1; }
        h = (unsigned char *)(((uintptr_t)m + (2u << 20) - 1) / (2u << 20) * (2u << 20));
        if (madvise(h, bytes, MADV_HUGEPAGE) != 0) printf("madvise(MADV_HUGEPAGE) failed\n");
        for (size_t i = 0; i < bytes; i += 4096) h[i] = (unsigned char)i;  // fault the pages in (as huge pages)
        CK(cudaHostRegister(h, bytes, cudaHostRegisterDefault));
        printf("THP-backed, cudaHostRegister'ed buffer\n");
    } else {
        CK(cudaHostAlloc(&h, P * 64, cudaHostAllocDefault));
        for (long long i = 0; i < P * 64; i += 4096) h[i] = (unsigned char)i;
    }
    std::vector<int> pix(n);
    srand(1);
    for (int v = 0; v < 10; ++v) {
        std::vector<int> p(n / 10);
        for (auto &x : p) x = (int)((((long long)rand() * 32768 + rand()) * 31 + rand()) % P);
        std::sort(p.begin(), p.end());
        std::copy(p.begin(), p.end(), pix.begin() + v * (n / 10));
    }
    int *d_pix; unsigned *d_dst;
    CK(cudaMalloc(&d_pix, n * 4)); CK(cudaMalloc(&d_dst, (size_t)n * 64));
    CK(cudaMemcpy(d_pix, pix.data(), n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("pinned buffer: %.2f GB (%d x 19.96 Mpx x 64 B)\n", P * 64 / 1e9, n_img);
    for (int row_bytes : {1, 40}) {
        for (int blocks : {6, 12, 24, 48, 96, 148, 148 * 4, 148 * 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                for (int it = 0; it < 5; ++it) k_rows<<<blocks, 256>>>(h, d_pix, n, row_bytes, d_dst);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
                if (rep == 1) printf("row %2d B, %5d CTAs: %7.3f ms per 250k rows = %6.1f M rows/s\n", row_bytes, blocks, ms, n / ms / 1e3);
            }
        }
    }
    return 0;
}
