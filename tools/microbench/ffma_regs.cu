// Issue rate of FFMA with three distinct REGISTER sources on B200 (sm_100a), next to FMUL / FADD and to the instruction mix
// of the contracted ("fast") PairHMM cell update.  Settles whether a three-register FFMA holds its issue port for two
// cycles (DESIGN.md section 4.4 of round 1 claimed so from a probe whose multiplier and addend were constant-bank
// operands).  All operands are loaded from global memory, so ptxas cannot fold them into immediates or constants; every
// chain is independent; 8 warps per SMSP; cycles per warp-instruction from clock64().
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -ftz=true -fmad=false -o ffma_regs ffma_regs.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

// KIND 0: x = x * a[i]                 (FMUL, two registers)
//      1: x = x + b[i]                 (FADD, two registers)
//      2: x = fma(x, a[i], b[i])       (FFMA, three distinct registers per instruction, 2 N + N live operands)
//      3: x = fma(x, a0, b0)           (FFMA, three registers, multiplier and addend shared by all chains)
//      4: x = fma(x, a[i], x2[i]); x2 = x2 * a[i]   (FMUL + FFMA pairs, the fast kernel's mix: t = m*p; y = fma(y, c, t))
//      5: the fast cell update itself on N independent "rows": 4 FMUL + 4 FFMA with the kernel's operand pattern
template <int KIND, int N>
__global__ void __launch_bounds__(256) probe(const float* __restrict__ in, float* out, int iters, long long* cyc)
{
    float x[N], y[N], z[N], a[N], b[N], c[N], d[N], e[N], w[N];
    #pragma unroll
    for (int i = 0; i < N; ++i) {
        const int o = (threadIdx.x + 32 * i) & 1023;
        x[i] = in[o]; y[i] = in[o + 1024]; z[i] = in[o + 2048]; a[i] = in[o + 3072]; b[i] = in[o + 4096];
        c[i] = in[o + 5120]; d[i] = in[o + 6144]; e[i] = in[o + 7168]; w[i] = in[o + 8192];
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            #pragma unroll
            for (int i = 0; i < N; ++i) {
                if (KIND == 0) x[i] = __fmul_rn(x[i], a[i]);
                if (KIND == 1) x[i] = __fadd_rn(x[i], b[i]);
                if (KIND == 2) x[i] = __fmaf_rn(x[i], a[i], b[i]);
                if (KIND == 3) x[i] = __fmaf_rn(x[i], a[0], b[0]);
                if (KIND == 4) { y[i] = __fmul_rn(y[i], a[i]); x[i] = __fmaf_rn(x[i], b[i], y[i]); }
                if (KIND == 5) {
                    // M' = (fma(Yd, g, fma(Xd, g, Md * mm))) * w ; Y' = fma(Y, c, M * my) ; X' = fma(Xu, c, Mu * mx)
                    const float md = x[(i + N - 1) % N], xd = y[(i + N - 1) % N], yd = z[(i + N - 1) % N];
                    const float t5 = __fmaf_rn(yd, b[i], __fmaf_rn(xd, b[i], __fmul_rn(md, a[i])));
                    const float mn = __fmul_rn(t5, w[i]);
                    const float yn = __fmaf_rn(z[i], c[i], __fmul_rn(x[i], d[i]));
                    const float xn = __fmaf_rn(y[(i + 1) % N], c[i], __fmul_rn(mn, e[i]));
                    x[i] = mn; z[i] = yn; y[i] = xn;
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    #pragma unroll
    for (int i = 0; i < N; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND, int N>
void run(const char* name, int per_iter, int nsm, const float* in)
{
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, probe<KIND, N>, 256, 0));
    if (bps > 4) bps = 4;
    const int blocks = nsm * bps, iters = 4000;
    float* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * blocks * 256)); CK(cudaMalloc(&cyc, sizeof(long long) * blocks));
    probe<KIND, N><<<blocks, 256>>>(in, out, 200, cyc);
    CK(cudaDeviceSynchronize());
    probe<KIND, N><<<blocks, 256>>>(in, out, iters, cyc);
    CK(cudaDeviceSynchronize());
    long long* h = (long long*)malloc(sizeof(long long) * blocks);
    CK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double c = 0; for (int i = 0; i < blocks; ++i) c += h[i];
    c /= blocks;
    const double warps_per_smsp = bps * 8 / 4.0;
    const double instr = warps_per_smsp * (double)iters * 4.0 * per_iter;          // warp-instructions per SMSP
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, probe<KIND, N>));
    printf("{\"probe\":\"%s\",\"chains\":%d,\"regs\":%d,\"warps_per_smsp\":%.0f,\"cycles_per_warp_instr\":%.4f}\n", name, N, fa.numRegs,
           warps_per_smsp, c / instr);
    free(h); CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\":\"%s\",\"sms\":%d}\n", p.name, p.multiProcessorCount);
    float* in; CK(cudaMalloc(&in, sizeof(float) * 10240));
    float* h = (float*)malloc(sizeof(float) * 10240);
    for (int i = 0; i < 10240; ++i) h[i] = 0.5f + (i % 97) * 1e-3f;
    CK(cudaMemcpy(in, h, sizeof(float) * 10240, cudaMemcpyHostToDevice));
    const int nsm = p.multiProcessorCount;
    run<0, 8>("fmul r,r", 8, nsm, in);
    run<1, 8>("fadd r,r", 8, nsm, in);
    run<2, 8>("ffma r,r,r (distinct per chain)", 8, nsm, in);
    run<2, 16>("ffma r,r,r (distinct per chain)", 16, nsm, in);
    run<3, 8>("ffma r,r,r (shared multiplier and addend)", 8, nsm, in);
    run<4, 8>("fmul + ffma pairs", 16, nsm, in);
    run<5, 8>("fast cell update, 8 rows (4 fmul + 4 ffma per row)", 64, nsm, in);
    run<5, 16>("fast cell update, 16 rows", 128, nsm, in);
    return 0;
}
