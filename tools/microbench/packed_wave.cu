// packed_wave.cu -- does the PairHMM steady loop get faster when two reads share a lane and every FP32 instruction
// is a packed FMUL2 / FADD2 (sm_100 f32x2 forms)?  A replica of pmm_forward_kernel's branch-free loop (same shuffles,
// same LDS.128 weight fetch, same byte stream load, same dependency structure) in two forms:
//   PACK = 1 : K rows per lane, scalar FMUL / FADD                 (the shipped kernel: K = 19, W = 8)
//   PACK = 2 : K rows per lane for each of two reads, FMUL2 / FADD2 (K = 10, W = 16: 160 rows for a 151-base read)
// Prints lane-cells per second and the share of the FP32 issue peak that 12 instructions per cell would need.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -ftz=true -fmad=false -o packed_wave packed_wave.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b);
    // ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even under -fmad=false; a product written as fma(x, y, +0) keeps
    // its own rounding (operands are non-negative, so the sign of a zero product cannot differ either)
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(0ull));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b);
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b), z = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 pmul2(float2 a, float2 b) {     // plain packed multiply (free to be contracted: fast mode)
    unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b);
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float fma2(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float pmul2(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float mul2(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add2(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float shfl_up(float v, int w) { return __shfl_up_sync(0xffffffffu, v, 1, w); }
__device__ __forceinline__ float2 shfl_up(float2 v, int w) { return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1, w), __shfl_up_sync(0xffffffffu, v.y, 1, w)); }

template <int PACK> struct Vec { using T = float; };
template <> struct Vec<2> { using T = float2; };

constexpr int kStream = 4096;

template <int K, int W, int PACK, bool FUSED>
__global__ void __launch_bounds__(128, 2) wave(const float* __restrict__ params, const unsigned char* __restrict__ stream,
                                               float* __restrict__ out, int steps)
{
    using T = typename Vec<PACK>::T;
    constexpr int E = K * PACK;                 // floats of weights per lane and class
    constexpr int KQ = (E + 3) / 4;
    constexpr int CLS = KQ * 32 * 4;
    extern __shared__ float4 smem4[];
    float* wtab = reinterpret_cast<float*>(smem4) + (threadIdx.x >> 5) * 5 * CLS;
    const int lane = threadIdx.x & 31, l = lane % W;
    float* wlane = wtab + lane * 4;
    T pMM[K], pG[K], pMX[K], pMY[K], pC[K], M[K], X[K], Y[K];
    const T* pp = reinterpret_cast<const T*>(params) + (size_t)(blockIdx.x * 128 + threadIdx.x) * K * 5;
    #pragma unroll
    for (int j = 0; j < K; ++j) {
        pMM[j] = pp[j * 5 + 0]; pG[j] = pp[j * 5 + 1]; pMX[j] = pp[j * 5 + 2]; pMY[j] = pp[j * 5 + 3]; pC[j] = pp[j * 5 + 4];
        M[j] = pMM[j]; X[j] = pG[j]; Y[j] = pMX[j];
    }
    for (int h = 0; h < 5; ++h)
        for (int q = 0; q < KQ * 4; ++q) wlane[h * CLS + (q / 4) * 128 + (q % 4)] = 0.9f + 0.01f * h + 0.001f * q;
    __syncwarp();
    T dM = M[0], dX = X[0], dY = Y[0], sM = M[0], sX = X[0];

    auto step = [&](unsigned e) {
        const T inM = shfl_up(M[K - 1], W), inX = shfl_up(X[K - 1], W), inY = shfl_up(Y[K - 1], W);
        const float* wp = wlane + e * CLS;
        float wf[KQ * 4];
        #pragma unroll
        for (int m = 0; m < KQ; ++m) {
            const float4 v = *reinterpret_cast<const float4*>(wp + m * 128);
            wf[m * 4] = v.x; wf[m * 4 + 1] = v.y; wf[m * 4 + 2] = v.z; wf[m * 4 + 3] = v.w;
        }
        T Mn[K], Xn[K], Yn[K];
        #pragma unroll
        for (int j = K - 1; j >= 0; --j) {
            const T md = j ? M[j - 1] : dM, xd = j ? X[j - 1] : dX, yd = j ? Y[j - 1] : dY;
            T w;
            if constexpr (PACK == 2) w = make_float2(wf[2 * j], wf[2 * j + 1]); else w = wf[j];
            if constexpr (FUSED) {       // the fast mode's contraction: FMUL + 2 FFMA + FMUL, FMUL + FFMA
                const T t5 = fma2(yd, pG[j], fma2(xd, pG[j], pmul2(md, pMM[j])));
                Mn[j] = pmul2(t5, w);
                Yn[j] = fma2(Y[j], pC[j], pmul2(M[j], pMY[j]));
            } else {
            const T t3 = add2(mul2(md, pMM[j]), mul2(xd, pG[j]));
            const T t5 = add2(t3, mul2(yd, pG[j]));
            Mn[j] = mul2(t5, w);
            Yn[j] = add2(mul2(M[j], pMY[j]), mul2(Y[j], pC[j]));
            }
        }
        #pragma unroll
        for (int j = 0; j < K; ++j) {
            const T mu = j ? Mn[j - 1] : inM, xu = j ? Xn[j - 1] : inX;
            if constexpr (FUSED) Xn[j] = fma2(xu, pC[j], pmul2(mu, pMX[j])); else Xn[j] = add2(mul2(mu, pMX[j]), mul2(xu, pC[j]));
        }
        #pragma unroll
        for (int j = 0; j < K; ++j) { M[j] = Mn[j]; X[j] = Xn[j]; Y[j] = Yn[j]; }
        sM = add2(sM, M[K - 1]); sX = add2(sX, X[K - 1]);
        dM = inM; dX = inX; dY = inY;
    };

    const unsigned char* q = stream + (blockIdx.x * 7 + l) % 64;
    unsigned e = q[0];
    #pragma unroll 1
    for (int t = 0; t + 4 <= steps; t += 4) {
        unsigned en[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) en[u] = q[u + 1];
        step(e); step(en[0]); step(en[1]); step(en[2]);
        e = en[3];
        q += 4;
        if (q - stream > kStream - 80) q -= kStream - 160;
    }
    float r;
    if constexpr (PACK == 2) r = sM.x + sM.y + sX.x + sX.y; else r = sM + sX;
    out[blockIdx.x * 128 + threadIdx.x] = r;
}

template <int K, int W, int PACK, bool FUSED = false> void run(const char* name, int nsm, double peak)
{
    constexpr int E = K * PACK, KQ = (E + 3) / 4, smem = 4 * 5 * KQ * 32 * 4 * (int)sizeof(float);
    const int blocks = nsm * 2, steps = 40000;
    float* params; unsigned char* stream; float* out;
    std::vector<float> hp((size_t)blocks * 128 * K * 5 * PACK);
    for (size_t i = 0; i < hp.size(); ++i) hp[i] = 0.2f + 0.6f * (float)((i * 2654435761u) % 1000) / 1000.0f;
    std::vector<unsigned char> hs(kStream + 64);
    for (size_t i = 0; i < hs.size(); ++i) hs[i] = (unsigned char)((i * 2654435761u >> 13) % 5);
    CK(cudaMalloc(&params, hp.size() * 4)); CK(cudaMemcpy(params, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&stream, hs.size())); CK(cudaMemcpy(stream, hs.data(), hs.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, (size_t)blocks * 128 * 4));
    auto kern = wave<K, W, PACK, FUSED>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<blocks, 128, smem>>>(params, stream, out, steps); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); kern<<<blocks, 128, smem>>>(params, stream, out, steps); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double cells = (double)blocks * 128 * K * PACK * steps;
    const double rate = cells / (best * 1e-3);
    printf("{\"bench\":\"%s\",\"K\":%d,\"W\":%d,\"pack\":%d,\"regs\":%d,\"ctas_per_sm\":%d,\"ms\":%.3f,\"lane_cells_per_s\":%.4e,\"tcups_at_full_rows\":%.3f,"
           "\"frac_of_fp32_issue_peak\":%.4f,\"cycles_per_step_at_1965\":%.1f}\n",
           name, K, W, PACK, fa.numRegs, occ, best, rate, rate * 1e-12, rate * (FUSED ? 8 : 12) / peak,
           best * 1e-3 * 1.965e9 / steps * 1.0);
    CK(cudaFree(params)); CK(cudaFree(stream)); CK(cudaFree(out));
}

__global__ void __launch_bounds__(256) probe(float* sink, int iters)
{
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = 1.0f + threadIdx.x * 1e-3f + i;
    const float m = 0.99999f, c = 1e-6f;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int rep = 0; rep < 64; ++rep) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = (rep & 1) ? __fadd_rn(x[i], c) : __fmul_rn(x[i], m);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    float* sink; CK(cudaMalloc(&sink, (size_t)nsm * 8 * 256 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe<<<nsm * 8, 256>>>(sink, 2000); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); probe<<<nsm * 8, 256>>>(sink, 2000); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double peak = (double)nsm * 8 * 256 * 2000 * 512 / (ms * 1e-3);
    printf("{\"device\":\"%s\",\"sms\":%d,\"fp32_issue_peak\":%.4e}\n", prop.name, nsm, peak);
    run<19, 8, 1>("scalar_K19_W8", nsm, peak);
    run<10, 16, 1>("scalar_K10_W16", nsm, peak);
    run<10, 16, 2>("packed_K10_W16", nsm, peak);
    run<12, 16, 2>("packed_K12_W16", nsm, peak);
    run<8, 32, 2>("packed_K8_W32", nsm, peak);
    // fast mode (contracted): fraction is against 8 instructions per cell
    run<19, 8, 1, true>("fused_scalar_K19_W8", nsm, peak);
    run<10, 16, 2, true>("fused_packed_K10_W16", nsm, peak);
    run<12, 16, 2, true>("fused_packed_K12_W16", nsm, peak);
    return 0;
}
