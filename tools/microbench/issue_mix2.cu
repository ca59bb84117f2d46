// Issue-slot microbenchmark, inline-PTX edition (the compiler cannot collapse the chains).
// Question: on B200, do ALU ops (LOP3 / SELP / SHFL) co-issue with scalar FP32 and with packed FFMA2/FADD2?
// Reports SMSP cycles per loop iteration from the event time and the SM clock (clock64/globaltimer).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix2 issue_mix2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ unsigned long long gtime(){ unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

#define FMUL(x,a)   asm volatile("mul.rn.ftz.f32 %0, %0, %1;" : "+f"(x) : "f"(a))
#define FADD(x,a)   asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x) : "f"(a))
#define FMA2Z(x,a)  asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(a), "l"(0ull))
#define FADD2(x,a)  asm volatile("add.rn.ftz.f32x2 %0, %0, %1;" : "+l"(x) : "l"(a))
#define LOP3(u,v,w) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(v), "r"(w))
#define SELP(u,v,w,c) asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; selp.b32 %0, %1, %2, p; }" : "+r"(u) : "r"(v), "r"(w), "r"(c))
#define LOPP_SEL(u,v,w,h,m) asm volatile("{ .reg .pred p; .reg .b32 t; and.b32 t, %3, %4; setp.ne.u32 p, t, 0; selp.b32 %0, %1, %2, p; }" : "+r"(u) : "r"(v), "r"(w), "r"(h), "r"(m))
#define SHFL(u)     asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(u))

// KIND: FP kind 0 = 32 scalar FMUL/iter, 1 = 16 FFMA2(rz)/iter, 2 = 16 FADD2, 3 = none, 4 = 16 FMUL + 16 FADD
// AK: 0 none, 1 LOP3, 2 SELP (invariant predicate), 3 AND+SETP+SELP (per-cell predicate), 4 SHFL ; NA = count per iter
template<int KIND, int AK, int NA>
__global__ void __launch_bounds__(256) mix(float* out, float a, unsigned m0, int iters, unsigned long long* stamps)
{
    float x[8]; unsigned long long y[8]; unsigned u[8];
    unsigned long long a2 = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a*0.9999f);
    #pragma unroll
    for (int i=0;i<8;++i){ x[i] = 1.0f + threadIdx.x*1e-3f + i; y[i] = a2 + i; u[i] = m0 + threadIdx.x*77u + i*13u; }
    unsigned c[8];
    #pragma unroll
    for (int i=0;i<8;++i) c[i] = (threadIdx.x >> i) & 1;
    unsigned long long g0 = gtime(); long long c0 = clock64();
    #pragma unroll 1
    for (int it=0; it<iters; ++it) {
        #pragma unroll
        for (int rep=0; rep<4; ++rep) {
            #pragma unroll
            for (int i=0;i<8;++i) {
                if (KIND==0) FMUL(x[i], a);
                if (KIND==1 && (rep&1)==0) FMA2Z(y[i], a2);
                if (KIND==2 && (rep&1)==0) FADD2(y[i], a2);
                if (KIND==4) { if (rep&1) FADD(x[i], a); else FMUL(x[i], a); }
                const int k = rep*8+i;
                if (k*NA/32 != (k+1)*NA/32) {   // spread NA ops evenly over the 32 slots
                    const int j = (k*NA/32) & 7;
                    if (AK==1) LOP3(u[j], u[(j+1)&7], u[(j+2)&7]);
                    if (AK==2) SELP(u[j], u[(j+1)&7], u[(j+2)&7], c[j]);
                    if (AK==3) LOPP_SEL(u[j], u[(j+1)&7], u[(j+2)&7], u[(j+3)&7], m0);
                    if (AK==4) SHFL(u[j]);
                }
            }
        }
    }
    long long c1 = clock64(); unsigned long long g1 = gtime();
    float s=0;
    #pragma unroll
    for (int i=0;i<8;++i) s += x[i] + __uint_as_float(u[i]) + (float)y[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
    if (threadIdx.x==0){ stamps[2*blockIdx.x]=c1-c0; stamps[2*blockIdx.x+1]=g1-g0; }
}

template<int KIND, int AK, int NA>
void run(const char* name, int nsm)
{
    int bps = 4, blocks = nsm*bps, threads = 256, iters = 20000;   // 8 warps per SMSP
    float* out; unsigned long long* st;
    CK(cudaMalloc(&out, sizeof(float)*blocks*threads));
    CK(cudaMalloc(&st, sizeof(unsigned long long)*2*blocks));
    mix<KIND,AK,NA><<<blocks,threads>>>(out, 0.999f, 0x12345u, 2000, st);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    mix<KIND,AK,NA><<<blocks,threads>>>(out, 0.999f, 0x12345u, iters, st);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long)*2*blocks);
    CK(cudaMemcpy(h, st, sizeof(unsigned long long)*2*blocks, cudaMemcpyDeviceToHost));
    double cyc=0, ns=0; for(int i=0;i<blocks;++i){ cyc+=h[2*i]; ns+=h[2*i+1]; }
    double mhz = cyc/ns*1e3;
    double cyc_iter = ms*1e-3*mhz*1e6/(8.0*iters);
    int nfp = KIND==0||KIND==4 ? 32 : (KIND==3 ? 0 : 16);
    int nal = AK==0 ? 0 : (AK==3 ? 3*NA : NA);
    printf("{\"mix\":\"%s\",\"fp_instr\":%d,\"other_instr\":%d,\"ms\":%.3f,\"sm_mhz\":%.0f,\"smsp_cyc_per_iter\":%.2f,\"ipc\":%.3f}\n",
        name, nfp, nal, ms, mhz, cyc_iter, (nfp+nal+3)/cyc_iter);
    free(h); CK(cudaFree(out)); CK(cudaFree(st));
}

int main(){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
    int nsm = p.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d}\n", p.name, nsm);
    run<0,0,0>("32fmul", nsm); run<4,0,0>("16fmul+16fadd", nsm); run<1,0,0>("16ffma2rz", nsm); run<2,0,0>("16fadd2", nsm);
    run<3,1,32>("32lop3", nsm); run<3,2,32>("32selp", nsm); run<3,3,16>("16x(and,setp,selp)", nsm); run<3,4,16>("16shfl", nsm);
    run<0,1,8>("32fmul+8lop3", nsm); run<0,1,16>("32fmul+16lop3", nsm); run<0,1,32>("32fmul+32lop3", nsm);
    run<0,2,8>("32fmul+8selp", nsm); run<0,2,16>("32fmul+16selp", nsm); run<0,2,32>("32fmul+32selp", nsm);
    run<0,3,4>("32fmul+4x(and,setp,selp)", nsm); run<0,3,8>("32fmul+8x(and,setp,selp)", nsm);
    run<1,1,8>("16ffma2rz+8lop3", nsm); run<1,1,16>("16ffma2rz+16lop3", nsm); run<1,1,32>("16ffma2rz+32lop3", nsm);
    run<1,2,8>("16ffma2rz+8selp", nsm); run<1,2,16>("16ffma2rz+16selp", nsm); run<1,2,32>("16ffma2rz+32selp", nsm);
    run<1,3,4>("16ffma2rz+4x(and,setp,selp)", nsm); run<1,3,8>("16ffma2rz+8x(and,setp,selp)", nsm);
    run<0,4,2>("32fmul+2shfl", nsm); run<0,4,4>("32fmul+4shfl", nsm); run<1,4,2>("16ffma2rz+2shfl", nsm); run<1,4,4>("16ffma2rz+4shfl", nsm);
    return 0;
}
