// Microbenchmark: per-SM throughput of the integer SIMD ops the epi8 kernel uses (warp-instructions / clk / SM).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#define ITER 4096
template<int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c){
	if(OP == 0) return __viaddmax_s16x2(a, b, c);
	if(OP == 1) return __vmaxs2(a, b);
	if(OP == 2) return __vimax3_s16x2(a, b, c);
	if(OP == 3) return __vadd2(a, b);
	if(OP == 4){ uint32_t d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
	if(OP == 5) return (a & b) ^ c;                 // LOP3
	if(OP == 6) return a * b + c;                   // IMAD
	if(OP == 7) return a + b + c;                   // IADD3
	if(OP == 8) return (uint32_t)max((int)(a + b), (int)c); // 32-bit VIADDMNMX
	if(OP == 9){ __half2 x = *(__half2*)&a, y = *(__half2*)&b; __half2 r = __hadd2(x, y); return *(uint32_t*)&r; }
	if(OP == 10){ __half2 x = *(__half2*)&a, y = *(__half2*)&b; __half2 r = __hmax2(x, y); return *(uint32_t*)&r; }
	if(OP == 11) return __viaddmin_s16x2_relu(a, b, c);
	if(OP == 12) return (uint32_t)__dp4a((int)a, (int)b, (int)c);
	if(OP == 13) return __vmaxu2(a, b);
	if(OP == 14) return (uint32_t)max(max((int)a, (int)b), (int)c); // 32-bit VIMNMX3
	if(OP == 15){ __half2 x = *(__half2*)&a, y = *(__half2*)&b, z = *(__half2*)&c; __half2 r = __hfma2_relu(x, y, z); return *(uint32_t*)&r; }
	if(OP == 16) return __funnelshift_r(a, b, c);
	return 0;
}
template<int OP> __global__ void k(uint32_t *out, uint32_t s0, uint32_t s1, long long *cyc){
	uint32_t r[8];
	for(int i=0;i<8;i++) r[i] = s0 + threadIdx.x * 17 + i;
	uint32_t b = s1 | 1, c = s0 ^ 0x00030003;
	long long t0 = clock64();
	for(int it=0;it<ITER;it++){
		#pragma unroll
		for(int i=0;i<8;i++) r[i] = op<OP>(r[i], r[(i + 1) & 7] | b, c);
	}
	long long t1 = clock64();
	uint32_t x = 0; for(int i=0;i<8;i++) x ^= r[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = x;
	if(threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<int OP> void run(const char *name, uint32_t *out, long long *cyc){
	int nb = 148 * 2, nt = 512; // 32 warps per SM
	k<OP><<<nb, nt>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
	k<OP><<<nb, nt>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
	long long h[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0; for(int i=0;i<nb;i++) avg += h[i]; avg /= nb;
	double winst_per_sm = 2.0 * (nt / 32) * 8.0 * ITER; // two CTAs per SM
	printf("%-28s %8.0f cycles  -> %.2f warp-inst/clk/SM (%.2f per SMSP)\n", name, avg, winst_per_sm / avg, winst_per_sm / avg / 4);
}
int main(){
	uint32_t *out; long long *cyc; cudaMalloc(&out, 296 * 512 * 4); cudaMalloc(&cyc, 296 * 8);
	run<0>("VIADDMNMX.S16x2", out, cyc); run<1>("VIMNMX.S16x2", out, cyc); run<2>("VIMNMX3.S16x2", out, cyc); run<3>("VIADD.16x2", out, cyc);
	run<4>("PRMT", out, cyc); run<5>("LOP3", out, cyc); run<6>("IMAD", out, cyc); run<7>("IADD3", out, cyc); run<8>("VIADDMNMX (32-bit)", out, cyc);
	run<9>("HADD2", out, cyc); run<10>("HMNMX2", out, cyc); run<11>("VIADDMNMX.S16x2.RELU(min)", out, cyc); run<12>("IDP.4A", out, cyc);
	run<13>("VIMNMX.U16x2", out, cyc); run<14>("VIMNMX3 (32-bit)", out, cyc); run<15>("HFMA2.RELU", out, cyc); run<16>("SHF (funnel)", out, cyc);
	return 0;
}
