// Issue-slot microbenchmark for B200 (sm_100a): does ALU/SHFL/LDS work co-issue with packed FP32 (FFMA2/FADD2)?
// Every chain is independent; one wave of CTAs; SM clock is derived from clock64()/globaltimer per CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -ftz=true -o issue_mix issue_mix.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ unsigned long long gtime(){ unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ float2 mul2x(float2 a, float2 b){ return __ffma2_rn(a, b, make_float2(0.f, 0.f)); }

// FPK: 0 scalar FMUL, 1 FFMA2(rz) (=exact packed mul), 2 FADD2, 3 scalar FADD, 4 FFMA scalar, 5 FFMA2 real
// ALK: 0 LOP3, 1 FSEL(pred from loop-invariant), 2 SHFL, 3 LDS, 4 IADD3, 5 MOV-ish (PRMT)
template<int FPK, int NFP, int ALK, int NAL>
__global__ void __launch_bounds__(256) mix(float* out, float a, float b, unsigned m0, int iters, unsigned long long* stamps)
{
    __shared__ float sm[256];
    sm[threadIdx.x] = a + threadIdx.x;
    __syncthreads();
    float2 x[NFP > 0 ? NFP : 1];
    unsigned u[NAL > 0 ? NAL : 1];
    float2 a2 = make_float2(a, a*1.0001f), b2 = make_float2(b, b*1.0001f);
    #pragma unroll
    for (int i=0;i<NFP;++i) x[i] = make_float2(1.0f + threadIdx.x*1e-3f + i, 2.0f+i);
    #pragma unroll
    for (int i=0;i<NAL;++i) u[i] = m0 + threadIdx.x*77u + i;
    unsigned long long g0 = gtime(); long long c0 = clock64();
    for (int it=0; it<iters; ++it) {
        #pragma unroll
        for (int rep=0; rep<4; ++rep) {
            #pragma unroll
            for (int i=0;i<(NFP>NAL?NFP:NAL);++i) {
                if (i<NFP) {
                    if (FPK==0) x[i].x = __fmul_rn(x[i].x, a);
                    if (FPK==1) x[i] = mul2x(x[i], a2);
                    if (FPK==2) x[i] = __fadd2_rn(x[i], b2);
                    if (FPK==3) x[i].x = __fadd_rn(x[i].x, b);
                    if (FPK==4) x[i].x = __fmaf_rn(x[i].x, a, b);
                    if (FPK==5) x[i] = __ffma2_rn(x[i], a2, b2);
                }
                if (i<NAL) {
                    if (ALK==0) u[i] = (u[i] & u[(i+1)%NAL]) ^ m0;
                    if (ALK==1) u[i] = ((threadIdx.x >> (i&7)) & 1) ? u[(i+1)%NAL] : u[(i+2)%NAL];
                    if (ALK==2) u[i] = __shfl_up_sync(0xffffffffu, u[i], 1);
                    if (ALK==3) u[i] = __float_as_uint(sm[u[i] & 255]);
                    if (ALK==4) u[i] = u[i] + u[(i+1)%NAL] + m0;
                    if (ALK==5) u[i] = __byte_perm(u[i], u[(i+1)%NAL], 0x5140);
                }
            }
        }
    }
    long long c1 = clock64(); unsigned long long g1 = gtime();
    float s=0;
    #pragma unroll
    for (int i=0;i<NFP;++i) s+=x[i].x+x[i].y;
    #pragma unroll
    for (int i=0;i<NAL;++i) s+=__uint_as_float(u[i]);
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
    if (threadIdx.x==0){ stamps[2*blockIdx.x]=c1-c0; stamps[2*blockIdx.x+1]=g1-g0; }
}

template<int FPK, int NFP, int ALK, int NAL>
void run(const char* name, int nsm)
{
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, mix<FPK,NFP,ALK,NAL>, 256, 0));
    if (bps > 4) bps = 4;               // 4 CTAs x 8 warps = 8 warps per SMSP
    int blocks = nsm*bps, threads = 256, iters = 20000;
    float* out; unsigned long long* st;
    CK(cudaMalloc(&out, sizeof(float)*blocks*threads));
    CK(cudaMalloc(&st, sizeof(unsigned long long)*2*blocks));
    mix<FPK,NFP,ALK,NAL><<<blocks,threads>>>(out, 0.999f, 1e-3f, 0x12345u, 2000, st);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    mix<FPK,NFP,ALK,NAL><<<blocks,threads>>>(out, 0.999f, 1e-3f, 0x12345u, iters, st);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long)*2*blocks);
    CK(cudaMemcpy(h, st, sizeof(unsigned long long)*2*blocks, cudaMemcpyDeviceToHost));
    double cyc=0, ns=0; for(int i=0;i<blocks;++i){ cyc+=h[2*i]; ns+=h[2*i+1]; }
    double mhz = cyc/ns*1e3; cyc/=blocks;
    double warps_per_smsp = bps*8/4.0;
    double fp_per_smsp = warps_per_smsp*iters*4.0*NFP, al_per_smsp = warps_per_smsp*iters*4.0*NAL;
    (void)fp_per_smsp; (void)al_per_smsp;
    printf("{\"mix\":\"%s\",\"tmpl\":\"%d,%d,%d,%d\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"sm_mhz\":%.0f,\"cyc_per_warp_iter\":%.3f}\n",
        name, FPK, NFP, ALK, NAL, bps, ms, mhz, cyc/(warps_per_smsp*iters));
    free(h); CK(cudaFree(out)); CK(cudaFree(st));
}

int main(){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
    int nsm = p.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d}\n", p.name, nsm);
    run<0,8,0,0>("fmul", nsm);       run<3,8,0,0>("fadd", nsm);      run<4,8,0,0>("ffma", nsm);
    run<1,8,0,0>("ffma2rz", nsm);    run<2,8,0,0>("fadd2", nsm);     run<5,8,0,0>("ffma2", nsm);
    run<0,0,0,8>("lop3", nsm);       run<0,0,1,8>("fsel", nsm);      run<0,0,2,8>("shfl", nsm);
    run<0,0,3,8>("lds", nsm);        run<0,0,4,8>("iadd3", nsm);     run<0,0,5,8>("prmt", nsm);
    run<0,8,0,2>("fmul+lop3", nsm);  run<0,8,0,4>("fmul+lop3", nsm); run<0,8,0,8>("fmul+lop3", nsm);
    run<1,8,0,2>("ffma2rz+lop3", nsm); run<1,8,0,4>("ffma2rz+lop3", nsm); run<1,8,0,8>("ffma2rz+lop3", nsm);
    run<2,8,0,4>("fadd2+lop3", nsm);
    run<1,8,1,4>("ffma2rz+fsel", nsm); run<1,8,1,8>("ffma2rz+fsel", nsm);
    run<0,8,1,4>("fmul+fsel", nsm);
    run<1,8,2,1>("ffma2rz+shfl", nsm); run<1,8,2,2>("ffma2rz+shfl", nsm); run<0,8,2,1>("fmul+shfl", nsm);
    run<1,8,3,1>("ffma2rz+lds", nsm);  run<1,8,3,2>("ffma2rz+lds", nsm);
    run<1,8,4,4>("ffma2rz+iadd3", nsm); run<1,8,5,4>("ffma2rz+prmt", nsm);
    return 0;
}
