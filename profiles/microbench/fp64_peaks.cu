// Microbenchmarks that fix the roofline denominators MEASURED_PEAKS.json does not carry:
//   * FP64 DFMA issue peak (vector pipe)
//   * FP64 DMMA peak for mma.sync m8n8k4 / m16n8k8 / m16n8k16 .f64
//   * DFMA and DMMA issued together (are they one pipe or two?)
//   * write-only HBM bandwidth for the store shapes the tabulation kernels emit
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double d[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d[i][0] = threadIdx.x; d[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(d[i][0], d[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
    if (s == 123.456) out[0] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double d[ILP][4]; double av[4] = {a, a + 1, a + 2, a + 3}; double bv[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d[i][0] = threadIdx.x; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma1688(d[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 123.456) out[0] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
    double d[ILP][4]; double av[8]; double bv[4];
    for (int i = 0; i < 8; ++i) av[i] = a + i;
    for (int i = 0; i < 4; ++i) bv[i] = b + i;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d[i][0] = threadIdx.x; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma16816(d[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 123.456) out[0] = s;
}
// DFMA and DMMA interleaved in every warp
template <int ILP>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    double d[ILP][2]; double acc[2 * ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d[i][0] = threadIdx.x; d[i][1] = i; acc[2 * i] = i; acc[2 * i + 1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            dmma884(d[i][0], d[i][1], a, b);
            acc[2 * i] = fma(acc[2 * i], a, b);
            acc[2 * i + 1] = fma(acc[2 * i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + acc[2 * i] + acc[2 * i + 1];
    if (s == 123.456) out[0] = s;
}
// DFMA with one operand from shared memory (broadcast) -- models coefficient fetch
template <int ILP>
__global__ void __launch_bounds__(256) k_dfma_lds(double* out, int iters, double b) {
    __shared__ double coef[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) coef[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
        double c = coef[it & 511];
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], c, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// ---- write-only bandwidth probes ----
// (a) flat: each thread writes consecutive doubles, grid-stride
__global__ void k_write_flat(double* out, size_t n, double v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = v;
}
__global__ void k_write_flat2(double2* out, size_t n2, double v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) out[i] = make_double2(v, v);
}
// (b) tabulation shape: out[row][p], thread = point, loops over rows (stride npts)
template <bool CS>
__global__ void k_write_rows(double* out, size_t npts, int nrows, double v) {
    size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= npts) return;
    for (int r = 0; r < nrows; ++r) {
        double val = v + r;
        if (CS) __stcs(out + (size_t)r * npts + p, val); else out[(size_t)r * npts + p] = val;
    }
}
// (c) DMMA accumulator shape: each warp-store writes 8 rows x 64 B
__global__ void k_write_mma(double* out, size_t npts, int nrows, double v) {
    // warp handles 8 points x all rows; lane>>2 = row-in-block, (lane&3)*2 = point pair
    size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    size_t p0 = warp * 8;
    if (p0 >= npts) return;
    for (int r = 0; r < nrows; r += 8) {
        int row = r + (lane >> 2);
        double2 val = make_double2(v + r, v);
        *reinterpret_cast<double2*>(out + (size_t)row * npts + p0 + (lane & 3) * 2) = val;
    }
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
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    double* out; size_t nbytes = (size_t)8 << 30; CK(cudaMalloc(&out, nbytes));
    const int iters = 4096;
    const int grid = sms * 8, block = 256;
    {
        float ms = time_ms([&] { k_dfma<16><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * grid * block * (double)iters * 16;
        printf("DFMA ilp16 occ8x256:   %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dfma<8><<<sms * 4, block>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * sms * 4 * block * (double)iters * 8;
        printf("DFMA ilp8  occ4x256:   %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dfma_lds<16><<<grid, block>>>(out, iters, 1e-9); });
        fl = 2.0 * grid * block * (double)iters * 16;
        printf("DFMA+LDS bcast 1:16:   %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dfma_lds<4><<<grid, block>>>(out, iters, 1e-9); });
        fl = 2.0 * grid * block * (double)iters * 4;
        printf("DFMA+LDS bcast 1:4:    %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { k_dmma884<16><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 8 * 8 * 4 * (double)grid * (block / 32) * iters * 16;
        printf("DMMA m8n8k4 ilp16:     %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma1688<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 16 * 8 * 8 * (double)grid * (block / 32) * iters * 8;
        printf("DMMA m16n8k8 ilp8:     %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 16 * 8 * 16 * (double)grid * (block / 32) * iters * 8;
        printf("DMMA m16n8k16 ilp8:    %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_mixed<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        double fl_mma = 2.0 * 8 * 8 * 4 * (double)grid * (block / 32) * iters * 8;
        double fl_fma = 2.0 * grid * block * (double)iters * 16;
        printf("mixed DMMA884+2DFMA:   %8.3f ms  %7.2f TFLOP/s (mma %5.2f + fma %5.2f)\n", ms, (fl_mma + fl_fma) / ms * 1e-9, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
    }
    {
        size_t n = nbytes / 8;
        float ms = time_ms([&] { k_write_flat<<<sms * 16, 512>>>(out, n, 1.5); });
        printf("write flat 8B:         %8.3f ms  %7.1f GB/s\n", ms, nbytes / ms * 1e-6);
        ms = time_ms([&] { k_write_flat2<<<sms * 16, 512>>>((double2*)out, n / 2, 1.5); });
        printf("write flat 16B:        %8.3f ms  %7.1f GB/s\n", ms, nbytes / ms * 1e-6);
        CK(cudaMemset(out, 0, 16));
        ms = time_ms([&] { CK(cudaMemsetAsync(out, 0, nbytes)); });
        printf("cudaMemset:            %8.3f ms  %7.1f GB/s\n", ms, nbytes / ms * 1e-6);
        for (int nrows : {30, 1650, 5324}) {
            size_t npts = (nbytes / 8 / nrows) & ~(size_t)255;
            double bytes = (double)npts * nrows * 8;
            for (int blk : {128, 256}) {
                ms = time_ms([&] { k_write_rows<false><<<(unsigned)((npts + blk - 1) / blk), blk>>>(out, npts, nrows, 2.5); });
                printf("write rows=%4d blk%3d:      %8.3f ms  %7.1f GB/s (npts %zu)\n", nrows, blk, ms, bytes / ms * 1e-6, npts);
                ms = time_ms([&] { k_write_rows<true><<<(unsigned)((npts + blk - 1) / blk), blk>>>(out, npts, nrows, 2.5); });
                printf("write rows=%4d blk%3d .cs:  %8.3f ms  %7.1f GB/s\n", nrows, blk, ms, bytes / ms * 1e-6);
            }
            if (nrows % 8 == 0 || nrows == 1650) {
                int nr = nrows & ~7;
                size_t warps = npts / 8;
                ms = time_ms([&] { k_write_mma<<<(unsigned)((warps * 32 + 255) / 256), 256>>>(out, npts, nr, 2.5); });
                printf("write mma-shape rows=%4d:   %8.3f ms  %7.1f GB/s\n", nr, ms, (double)npts * nr * 8 / ms * 1e-6);
            }
        }
    }
    // sustained DFMA for ~2 s to see clocks under FP64 load
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int launches = 60;
        for (int i = 0; i < launches; ++i) k_dfma<16><<<grid, block>>>(out, iters * 4, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * grid * block * (double)iters * 4 * 16 * launches;
        printf("DFMA sustained %.1f s:  %7.2f TFLOP/s\n", ms * 1e-3, fl / ms * 1e-9);
    }
    CK(cudaFree(out));
    return 0;
}
