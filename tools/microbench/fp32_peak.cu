// FP32 issue/pipe microbenchmark for B200 (sm_100a).
// Measures the denominators the PairHMM roofline is quoted against (SURVEY.md §8d says the FP32 peak
// must be measured on the box, MEASURED_PEAKS.json has none), and probes whether the packed
// FMUL2/FADD2/FFMA2 forms free issue slots for ALU/SHFL/LDS work next to the FMA pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -ftz=true -o fp32_peak fp32_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <cmath>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int ITERS = 4096;
constexpr int NCH = 8;   // independent chains per thread

// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it does not for the scalar forms), which
// changes rounding.  A product written as fma(x, y, +0) keeps its own rounding and is left alone.
__device__ __forceinline__ float2 mul2x(float2 a, float2 b){ return __ffma2_rn(a, b, make_float2(0.f, 0.f)); }

// mode 0: scalar FMUL ; 1: scalar FMUL+FADD alternating ; 2: scalar FFMA
// mode 3: FMUL2 ; 4: FMUL2+FADD2 ; 5: FFMA2
// mode 6: scalar 12 FP (8 FMUL + 4 FADD) + 2 ALU (LOP3-pred + FSEL) per "cell"
// mode 7: packed 12 FP2 + 4 ALU per 2 cells
// mode 8: packed 12 FP2 + 4 ALU + shfl every 8 cells
// mode 9: SHFL only ; 10: DMUL+DADD ; 11: DFMA
template<int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float a0, float b0, unsigned m0, long long* cyc)
{
    float a = a0, b = b0;
    long long t0 = clock64();
    if (MODE <= 2) {
        float x[NCH];
        #pragma unroll
        for (int i=0;i<NCH;++i) x[i] = 1.0f + threadIdx.x*1e-3f + i;
        for (int it=0; it<ITERS; ++it) {
            #pragma unroll
            for (int i=0;i<NCH;++i) {
                if (MODE==0) x[i] = __fmul_rn(x[i], a);
                if (MODE==1) { x[i] = __fmul_rn(x[i], a); x[i] = __fadd_rn(x[i], b); }
                if (MODE==2) x[i] = __fmaf_rn(x[i], a, b);
            }
        }
        float s=0; 
        #pragma unroll
        for (int i=0;i<NCH;++i) s+=x[i];
        out[blockIdx.x*blockDim.x+threadIdx.x]=s;
    } else if (MODE <= 5) {
        float2 x[NCH]; float2 a2 = make_float2(a, a*1.0001f), b2 = make_float2(b, b*1.0001f);
        #pragma unroll
        for (int i=0;i<NCH;++i) x[i] = make_float2(1.0f + threadIdx.x*1e-3f + i, 2.0f+i);
        for (int it=0; it<ITERS; ++it) {
            #pragma unroll
            for (int i=0;i<NCH;++i) {
                if (MODE==3) x[i] = __fmul2_rn(x[i], a2);
                if (MODE==4) { x[i] = mul2x(x[i], a2); x[i] = __fadd2_rn(x[i], b2); }
                if (MODE==5) x[i] = __ffma2_rn(x[i], a2, b2);
            }
        }
        float s=0;
        #pragma unroll
        for (int i=0;i<NCH;++i) s+=x[i].x+x[i].y;
        out[blockIdx.x*blockDim.x+threadIdx.x]=s;
    } else if (MODE == 6) {
        // scalar cell: M = ((Md*pMM + Xd*pG) + Yd*pG) * w ; X = Mu*pMX + Xu*pXX ; Y = Ml*pMY + Yl*pYY
        float M[NCH], X[NCH], Y[NCH], P[NCH][7];
        #pragma unroll
        for (int i=0;i<NCH;++i){ M[i]=1.0f+i; X[i]=0.5f+i; Y[i]=0.25f+i+threadIdx.x;
            #pragma unroll
            for (int q=0;q<7;++q) P[i][q] = out[(threadIdx.x*NCH+i)*7+q]; }
        unsigned hm = m0 + threadIdx.x*2654435761u;
        for (int it=0; it<ITERS/4; ++it) {
            hm = hm*1664525u + 1013904223u;
            float Mu = M[0], Xu = X[0];
            #pragma unroll
            for (int i=NCH-1;i>=1;--i) {
                bool p = (hm & (0x1u<<i)) != 0;
                float w = p ? P[i][5] : P[i][6];
                float t = __fadd_rn(__fadd_rn(__fmul_rn(M[i-1],P[i][0]), __fmul_rn(X[i-1],P[i][1])), __fmul_rn(Y[i-1],P[i][1]));
                float Mn = __fmul_rn(t, w);
                float Yn = __fadd_rn(__fmul_rn(M[i],P[i][2]), __fmul_rn(Y[i],P[i][3]));
                float Xn = __fadd_rn(__fmul_rn(Mu,P[i][4]), __fmul_rn(Xu,P[i][3]));
                M[i]=Mn; X[i]=Xn; Y[i]=Yn; Mu = Mn; Xu = Xn;
            }
            M[0] = Mu; X[0] = Xu; Y[0] = Y[NCH-1];
        }
        float s=0;
        #pragma unroll
        for (int i=0;i<NCH;++i) s+=M[i]+X[i]+Y[i];
        out[blockIdx.x*blockDim.x+threadIdx.x]=s;
    } else if (MODE == 7 || MODE == 8 || MODE == 12) {
        float2 M[NCH], X[NCH], Y[NCH], P[NCH][7];
        #pragma unroll
        for (int i=0;i<NCH;++i){ M[i]=make_float2(1.0f+i,2.f); X[i]=make_float2(0.5f+i,1.f); Y[i]=make_float2(0.25f+i+threadIdx.x,3.f);
            #pragma unroll
            for (int q=0;q<7;++q) P[i][q] = ((float2*)out)[(threadIdx.x*NCH+i)*7+q]; }
        unsigned hm = m0 + threadIdx.x*2654435761u;
        for (int it=0; it<ITERS/4; ++it) {
            hm = hm*1664525u + 1013904223u;
            float2 Mu = M[0], Xu = X[0];
            if (MODE==8) {
                Mu.x = __shfl_up_sync(0xffffffffu, M[NCH-1].x, 1); Mu.y = __shfl_up_sync(0xffffffffu, M[NCH-1].y, 1);
                Xu.x = __shfl_up_sync(0xffffffffu, X[NCH-1].x, 1); Xu.y = __shfl_up_sync(0xffffffffu, X[NCH-1].y, 1);
                Y[0].x = __shfl_up_sync(0xffffffffu, Y[NCH-1].x, 1); Y[0].y = __shfl_up_sync(0xffffffffu, Y[NCH-1].y, 1);
            }
            #pragma unroll
            for (int i=NCH-1;i>=1;--i) {
                bool p0 = (hm & (0x1u<<i)) != 0;
                bool p1 = (hm & (0x100u<<i)) != 0;
                float2 w = make_float2(p0 ? P[i][5].x : P[i][6].x, p1 ? P[i][5].y : P[i][6].y);
                float2 Mn, Xn, Yn;
                if (MODE != 12) {
                    float2 t = __fadd2_rn(__fadd2_rn(mul2x(M[i-1],P[i][0]), mul2x(X[i-1],P[i][1])), mul2x(Y[i-1],P[i][1]));
                    Mn = mul2x(t, w);
                    Yn = __fadd2_rn(mul2x(M[i],P[i][2]), mul2x(Y[i],P[i][3]));
                    Xn = __fadd2_rn(mul2x(Mu,P[i][4]), mul2x(Xu,P[i][3]));
                } else {
                    float2 t = __ffma2_rn(M[i-1], P[i][0], __fmul2_rn(__fadd2_rn(X[i-1], Y[i-1]), P[i][1]));
                    Mn = __fmul2_rn(t, w);
                    Yn = __ffma2_rn(M[i], P[i][2], __fmul2_rn(Y[i], P[i][3]));
                    Xn = __ffma2_rn(Mu, P[i][4], __fmul2_rn(Xu, P[i][3]));
                }
                M[i]=Mn; X[i]=Xn; Y[i]=Yn; Mu = Mn; Xu = Xn;
            }
            M[0] = Mu; X[0] = Xu; Y[0] = Y[NCH-1];
        }
        float s=0;
        #pragma unroll
        for (int i=0;i<NCH;++i) s+=M[i].x+X[i].x+Y[i].x+M[i].y+X[i].y+Y[i].y;
        out[blockIdx.x*blockDim.x+threadIdx.x]=s;
    } else if (MODE == 9) {
        float x[NCH];
        #pragma unroll
        for (int i=0;i<NCH;++i) x[i] = 1.0f + threadIdx.x + i;
        for (int it=0; it<ITERS; ++it) {
            #pragma unroll
            for (int i=0;i<NCH;++i) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
        }
        float s=0;
        #pragma unroll
        for (int i=0;i<NCH;++i) s+=x[i];
        out[blockIdx.x*blockDim.x+threadIdx.x]=s;
    } else {
        double x[NCH]; double da=a, db=b;
        #pragma unroll
        for (int i=0;i<NCH;++i) x[i] = 1.0 + threadIdx.x*1e-3 + i;
        for (int it=0; it<ITERS/4; ++it) {
            #pragma unroll
            for (int i=0;i<NCH;++i) {
                if (MODE==10) { x[i] = __dmul_rn(x[i], da); x[i] = __dadd_rn(x[i], db); }
                if (MODE==11) x[i] = __fma_rn(x[i], da, db);
            }
        }
        double s=0;
        #pragma unroll
        for (int i=0;i<NCH;++i) s+=x[i];
        out[blockIdx.x*blockDim.x+threadIdx.x]=(float)s;
    }
    long long t1 = clock64();
    if (threadIdx.x==0) cyc[blockIdx.x] = t1-t0;
}

// FTZ tininess probe: products whose exact value is just below FLT_MIN.
__global__ void ftz_probe(const float* x, const float* y, float* o, int n){
    int i = threadIdx.x; if (i<n) { o[i] = __fmul_rn(x[i], y[i]); o[n+i] = __fadd_rn(x[i], -y[i]); }
}

template<int MODE> void run(const char* name, double lane_ops_per_thread_iter, double issue_per_thread_iter, int iters, int nsm, double* out_rate=nullptr)
{
    int blocks = nsm*8, threads = 256;
    float* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float)*(blocks*threads + 256*NCH*7*2))); CK(cudaMemset(out, 0x3c, sizeof(float)*(blocks*threads + 256*NCH*7*2)));
    CK(cudaMalloc(&cyc, sizeof(long long)*blocks));
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<MODE><<<blocks,threads>>>(out, 0.999f, 1e-3f, 0x12345u, cyc); // warm-up
    CK(cudaDeviceSynchronize());
    float best=1e30f;
    for (int r=0;r<5;++r){
        CK(cudaEventRecord(e0));
        bench<MODE><<<blocks,threads>>>(out, 0.999f, 1e-3f, 0x12345u, cyc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if (ms<best) best=ms;
    }
    long long* h = (long long*)malloc(sizeof(long long)*blocks);
    CK(cudaMemcpy(h, cyc, sizeof(long long)*blocks, cudaMemcpyDeviceToHost));
    double avg=0; for(int i=0;i<blocks;++i) avg+=h[i]; avg/=blocks;
    double total_lane_ops = (double)blocks*threads*iters*lane_ops_per_thread_iter;
    double total_issue = (double)blocks*(threads/32)*iters*issue_per_thread_iter;
    double sec = best*1e-3;
    // per-SM per-clock figures from the block cycle counter: 8 blocks x 8 warps resident per SM
    double laneops_per_clk_sm = (double)8*threads*iters*lane_ops_per_thread_iter/avg;
    double issue_per_clk_sm = (double)8*(threads/32)*iters*issue_per_thread_iter/avg;
    printf("{\"bench\":\"%s\",\"ms\":%.4f,\"fp_lane_ops_per_s\":%.4e,\"warp_instr_per_s\":%.4e,\"fp_lane_ops_per_clk_sm\":%.2f,\"warp_instr_per_clk_sm\":%.3f,\"eff_mhz\":%.0f}\n",
        name, best, total_lane_ops/sec, total_issue/sec, laneops_per_clk_sm, issue_per_clk_sm, avg/sec*1e-6);
    if (out_rate) *out_rate = total_lane_ops/sec;
    free(h); CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main(){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
    int nsm = p.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, nsm, p.clockRate);
    const int C = NCH-1;
    run<0>("fmul_scalar", NCH, NCH, ITERS, nsm);
    run<1>("fmul_fadd_scalar", 2*NCH, 2*NCH, ITERS, nsm);
    run<2>("ffma_scalar", NCH, NCH, ITERS, nsm);
    run<3>("fmul2", 2*NCH, NCH, ITERS, nsm);
    run<4>("fmul2_fadd2", 4*NCH, 2*NCH, ITERS, nsm);
    run<5>("ffma2", 2*NCH, NCH, ITERS, nsm);
    run<6>("cell_scalar_12fp_2alu", 12.0*C, 14.0*C+1, ITERS/4, nsm);
    run<7>("cell_packed_12fp2_4alu", 24.0*C, 16.0*C+1, ITERS/4, nsm);
    run<8>("cell_packed_shfl", 24.0*C, 16.0*C+7, ITERS/4, nsm);
    run<12>("cell_packed_fused_8fp2_4alu", 24.0*C, 12.0*C+1, ITERS/4, nsm);
    run<9>("shfl", 0, NCH, ITERS, nsm);
    run<10>("dmul_dadd", 2*NCH, 2*NCH, ITERS/4, nsm);
    run<11>("dfma", NCH, NCH, ITERS/4, nsm);

    // FTZ tininess probe
    const int n = 8;
    float hx[n], hy[n], ho[2*n];
    // FLT_MIN * (1 - 2^-24 .. ) built as products: x = FLT_MIN*2^k, y = (1-eps)*2^-k
    float eps[n] = {ldexpf(1.f,-24), ldexpf(1.f,-23), ldexpf(1.f,-25)*3, ldexpf(1.f,-22), 0.f, ldexpf(1.f,-20), ldexpf(1.f,-10), 0.5f};
    for (int i=0;i<n;++i){ hx[i] = ldexpf(1.0f + ldexpf(1.f,-23), -126+20) ; hy[i] = ldexpf(1.0f - eps[i], -20); }
    // exact product = 2^-126 * (1+2^-23)(1-eps): for eps=2^-23 => 2^-126*(1-2^-46) -> rounds to 2^-126 with unbounded exponent (not tiny after rounding)
    float *dx,*dy,*dout; CK(cudaMalloc(&dx,sizeof(hx))); CK(cudaMalloc(&dy,sizeof(hy))); CK(cudaMalloc(&dout,sizeof(ho)));
    CK(cudaMemcpy(dx,hx,sizeof(hx),cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy,hy,sizeof(hy),cudaMemcpyHostToDevice));
    ftz_probe<<<1,32>>>(dx,dy,dout,n); CK(cudaMemcpy(ho,dout,sizeof(ho),cudaMemcpyDeviceToHost));
    for (int i=0;i<n;++i){ unsigned u; memcpy(&u,&ho[i],4); unsigned ux,uy; memcpy(&ux,&hx[i],4); memcpy(&uy,&hy[i],4);
        printf("{\"ftz_probe\":%d,\"x\":\"0x%08x\",\"y\":\"0x%08x\",\"gpu_prod\":\"0x%08x\"}\n", i, ux, uy, u); }
    return 0;
}
