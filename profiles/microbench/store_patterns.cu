// Write-only bandwidth of the store shapes a DMMA epilogue can produce.
// out[row][p] (row stride = npts doubles).  A warp owns an 8-row x 32-point tile (8 x 256 B) and
// writes it with 16-byte stores arranged as R rows x (512/R) bytes per instruction.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// RPI = rows per instruction (8, 4, 2, 1); each instruction stores 32 lanes x 16 B = 512 B
template <int RPI>
__global__ void k_store(double* out, size_t npts, int nrows, double v) {
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t p0 = warp * 32;                 // 32 points per warp
    if (p0 >= npts) return;
    constexpr int LPR = 32 / RPI;                // lanes per row in one instruction
    const int r_in = lane / LPR, c_in = lane % LPR;
    for (int r = 0; r < nrows; r += 8) {
        // tile of 8 rows x 32 points = 8 x 16 double2; instruction i covers rows [i*RPI, i*RPI+RPI) x chunk
        constexpr int NINST = 8 * 16 / 32;       // 4 instructions of 32 double2 each... per 8x32 tile: 128 double2 -> 4 instr
#pragma unroll
        for (int i = 0; i < NINST; ++i) {
            // flatten: unit u = i*32 + lane; row-major within (RPI rows x LPR units) blocks
            const int blk = i;                    // block of RPI rows x (LPR double2) columns
            const int blocks_per_rowgroup = 16 / LPR;     // column blocks to cover 16 double2 of a row
            const int rg = blk / blocks_per_rowgroup, cb = blk % blocks_per_rowgroup;
            const int row = r + rg * RPI + r_in;
            const int col2 = cb * LPR + c_in;     // double2 index within the 32-point row segment
            if (row < nrows)
                *reinterpret_cast<double2*>(out + (size_t)row * npts + p0 + 2 * col2) = make_double2(v + row, v);
        }
    }
}
// variant with 8-byte stores, 1 row x 256 B per instruction (the thread-per-point kernels)
__global__ void k_store_rows(double* out, size_t npts, int nrows, double v) {
    size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= npts) return;
    for (int r = 0; r < nrows; ++r) out[(size_t)r * npts + p] = v + r;
}
// 2 points per thread: 1 row x 512 B per instruction with 16-byte stores
__global__ void k_store_rows2(double* out, size_t npts, int nrows, double v) {
    size_t p = 2 * (blockIdx.x * (size_t)blockDim.x + threadIdx.x);
    if (p >= npts) return;
    for (int r = 0; r < nrows; ++r) *reinterpret_cast<double2*>(out + (size_t)r * npts + p) = make_double2(v + r, v);
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    double* out; size_t nbytes = (size_t)8 << 30; CK(cudaMalloc(&out, nbytes));
    for (int nrows : {1648, 1264, 5320, 72}) {
        size_t npts = (nbytes / 8 / nrows) & ~(size_t)1023;
        double bytes = (double)npts * nrows * 8;
        unsigned gridw = (unsigned)((npts / 32 * 32 + 255) / 256);
        float ms;
        ms = time_ms([&] { k_store<8><<<gridw, 256>>>(out, npts, nrows, 1.0); });
        printf("rows %5d  8 rows x  64 B / instr: %7.1f GB/s\n", nrows, bytes / ms * 1e-6);
        ms = time_ms([&] { k_store<4><<<gridw, 256>>>(out, npts, nrows, 1.0); });
        printf("rows %5d  4 rows x 128 B / instr: %7.1f GB/s\n", nrows, bytes / ms * 1e-6);
        ms = time_ms([&] { k_store<2><<<gridw, 256>>>(out, npts, nrows, 1.0); });
        printf("rows %5d  2 rows x 256 B / instr: %7.1f GB/s\n", nrows, bytes / ms * 1e-6);
        ms = time_ms([&] { k_store<1><<<gridw, 256>>>(out, npts, nrows, 1.0); });
        printf("rows %5d  1 row  x 512 B / instr: %7.1f GB/s\n", nrows, bytes / ms * 1e-6);
        ms = time_ms([&] { k_store_rows<<<(unsigned)((npts + 127) / 128), 128>>>(out, npts, nrows, 1.0); });
        printf("rows %5d  1 row  x 256 B (8 B/thread): %7.1f GB/s\n", nrows, bytes / ms * 1e-6);
        ms = time_ms([&] { k_store_rows2<<<(unsigned)((npts / 2 + 127) / 128), 128>>>(out, npts, nrows, 1.0); });
        printf("rows %5d  1 row  x 512 B (16 B/thread, 2 pts): %7.1f GB/s\n", nrows, bytes / ms * 1e-6);
    }
    CK(cudaFree(out));
    return 0;
}
