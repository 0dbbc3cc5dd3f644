// epi8_kernels.cuh -- sm_100a kernels for the 8-bit banded DP (replaces bsalign.h:2094-4050 of the reference).
//
// Parallel decomposition:
//   * the reference's SSE word has 16 int8 lanes; lane j walks the "running block" of band positions
//     [j*W, (j+1)*W) sequentially, twice per row (pass 1: block-exit F; pass 2: the row).  To stay
//     bit-exact under int8 saturation we keep exactly that dependency structure: one GROUP of 8 threads
//     per pair, each thread owning two SSE lanes packed as s16x2 in one register so the native
//     VIADD.16x2 / VIMNMX(3).S16x2 / VIADDMNMX.S16x2 instructions do two cells per issue, with explicit
//     clamps to [-128,127] wherever the SSE code saturates.
//   * the previous row lives in shared memory, indexed by ABSOLUTE query position modulo the band width,
//     so the adaptive band shift (row_movx, bsalign.h:2244) is an index offset plus the few synthesized
//     overhang cells instead of a data shuffle; rows are updated in place.
//   * four groups share a warp and run the row loop in lock step (no divergence in the hot loops);
//     groups fetch new pairs from an atomic counter (persistent scheduling), longest pairs first.
//   * every finished row is streamed to the HBM traceback store with 16-byte coalesced stores
//     ((pw+1) bytes per cell + 80 bytes of row anchors, the reference's own trace format in band-circular
//     order), which is the kernel's algorithmic traffic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsb200 {

constexpr int kGroup = 8;            // threads per pair
constexpr int kLanes = 16;           // SSE lanes of the reference build (WORDSIZE, bsalign.h:142)
constexpr int kFwdThreads = 128;     // 16 groups per CTA
constexpr int kMetaInts = 20;        // per-row anchors: ub[17], rbeg, 2 pad  (80 B, bsalign.h:3878)
constexpr int kScoreMin = -536870911;
constexpr int kEpi8Min = -63;
constexpr int kEpi8Max = 63;

struct Epi8Args {
	const uint8_t *seqs;
	const uint64_t *qoff, *toff;
	const uint32_t *qlen, *tlen;
	const uint32_t *order;       // pair indices of this wave, heaviest first
	uint32_t npairs;             // pairs in this wave
	unsigned int *counter;       // work-stealing counter (zeroed before launch)
	uint8_t *trace;              // traceback arena
	const uint64_t *trace_off;   // per pair: byte offset of its block in the arena
	int32_t *results;            // per pair 10 ints; forward writes score/qe/te
	int32_t *status;             // per pair flags
	uint32_t bandwidth;          // requested (0 = full)
	uint32_t max_bw;             // largest rounded band in the batch (smem sizing)
	uint32_t group_smem;         // bytes of shared memory per group
	int mode;
	int8_t mtx[16];
	int8_t go1, ge1, go2, ge2;
	int8_t smax, smin;
};

// ---- s16x2 helpers: two int8 lanes per register, exact SSE saturation semantics ---------------------
__device__ __forceinline__ uint32_t pk(int lo, int hi){ return (uint32_t)(lo & 0xffff) | ((uint32_t)hi << 16); }
__device__ __forceinline__ uint32_t pk1(int v){ return pk(v, v); }
// bytes (zero-extended in a,b) -> sign-extended s16x2
// (raw PRMT: selector nibble 8|i replicates the sign of byte i; __byte_perm() masks that bit away)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel){ uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; }
__device__ __forceinline__ uint32_t pk_s8(uint32_t a, uint32_t b){ return prmt(a, b, 0xC480u); }
__device__ __forceinline__ uint32_t sadd(uint32_t a, uint32_t b){ return __vmins2(__viaddmax_s16x2(a, b, 0xff80ff80u), 0x007f007fu); }
__device__ __forceinline__ uint32_t ssub(uint32_t a, uint32_t b){ return sadd(a, __vneg2(b)); }
__device__ __forceinline__ uint32_t smax(uint32_t a, uint32_t b){ return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t smax3(uint32_t a, uint32_t b, uint32_t c){ return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ int lo16(uint32_t v){ return (int)(short)(v & 0xffff); }
__device__ __forceinline__ int hi16(uint32_t v){ return (int)(short)(v >> 16); }
__device__ __forceinline__ int clamp8(int v){ return max(-128, min(127, v)); }

__host__ __device__ __forceinline__ int epi8_piecewise(int go1, int ge1, int go2, int ge2, int bw){
	if(go2 < go1 && ge2 > ge1 && go2 + ge2 < go1 + ge1 && (go1 - go2) / (ge1 - ge2) < bw) return 2; // bsalign.h:2084-2092
	return go1 ? 1 : 0;
}

// sum of `count` u cells starting at circular slot s0, spread over the 8 threads of a group
// (callers sit in branches that only the 8 threads of one group take together, so the shuffles name exactly those lanes)
__device__ __forceinline__ int group_sum(const int8_t *sU, uint32_t bw, uint32_t s0, uint32_t count, int t){
	const unsigned gm = 0xffu << ((threadIdx.x & 31) & 24);
	int s = 0;
	for(uint32_t k=t;k<count;k+=kGroup){
		uint32_t sl = s0 + k; if(sl >= bw) sl -= bw;
		s += sU[sl];
	}
	s += __shfl_xor_sync(gm, s, 1);
	s += __shfl_xor_sync(gm, s, 2);
	s += __shfl_xor_sync(gm, s, 4);
	return s;
}

// absolute H at band position pos (bsalign.h:3187-3197); pos must be inside [0,bw)
__device__ __forceinline__ int group_getscore(const int8_t *sU, const int32_t *sUB, uint32_t bw, uint32_t W, uint32_t rslot, uint32_t pos, int t){
	uint32_t j = pos / W, i = pos - j * W;
	uint32_t s0 = rslot + j * W; if(s0 >= bw) s0 -= bw;
	return sUB[j] + group_sum(sU, bw, s0, i + 1, t);
}

template<int PW>
__global__ void __launch_bounds__(kFwdThreads) epi8_forward_kernel(const Epi8Args a){
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int lane = threadIdx.x & 31;
	const int t = lane & 7;
	const int A = 2 * t, B = A + 1;
	uint8_t *gs = smem_raw + (size_t)(threadIdx.x >> 3) * a.group_smem;
	int8_t *sU = (int8_t*)gs;
	int8_t *sE = sU + a.max_bw;
	int8_t *sQ = sE + (PW >= 1 ? a.max_bw : 0);
	uint8_t *sC = (uint8_t*)(sQ + (PW == 2 ? a.max_bw : 0));
	int32_t *sUB = (int32_t*)(sC + a.max_bw);      // kMetaInts ints: ub[17], rbeg
	int32_t *sTmp = sUB + kMetaInts;               // 20 ints scratch
	int8_t *sF = (int8_t*)(sTmp + kMetaInts);      // 16 fend, 16 gend
	int32_t *sRM = (int32_t*)(sF + 32);            // 32 ints scratch for row_max

	const int mode = a.mode & 3;
	const int go1 = a.go1, ge1 = a.ge1, go2 = a.go2, ge2 = a.ge2;
	const int GOEi = (int8_t)(go1 + ge1), GQPi = (int8_t)(go2 + ge2);
	const uint32_t GE = pk1(ge1), GOE = pk1(GOEi), GP = pk1(ge2), GQP = pk1(GQPi);
	const uint32_t GOQ = pk1(clamp8(GOEi - GQPi));
	const uint32_t NGOE = pk1(-GOEi), NGOQ = pk1(-clamp8(GOEi - GQPi)), NGQP = pk1(-GQPi);
	// matrix columns: colw[tb] holds mtx[0*4+tb], mtx[1*4+tb], mtx[2*4+tb], mtx[3*4+tb] as bytes
	uint32_t colw[4];
	#pragma unroll
	for(int c=0;c<4;c++) colw[c] = (uint32_t)(uint8_t)a.mtx[c] | ((uint32_t)(uint8_t)a.mtx[4 + c] << 8) | ((uint32_t)(uint8_t)a.mtx[8 + c] << 16) | ((uint32_t)(uint8_t)a.mtx[12 + c] << 24);

	bool have = false, done = false;
	uint32_t pair = 0, qlen = 1, tlen = 1, bw = 16, W = 1, row = 0, rbeg = 0, rslot = 0, mov = 0;
	const uint8_t *qs = a.seqs, *ts = a.seqs;
	uint8_t *tr = a.trace;       // this pair's trace block (row -1 first)
	int32_t *meta = nullptr;     // this pair's anchors block
	uint32_t RS = 16;            // bytes per trace row
	int best = kScoreMin, best_qe = 0, best_te = 0, stflag = 0;

	while(true){
		if(!have && !done){
			uint32_t idx = 0;
			if(t == 0) idx = atomicAdd(a.counter, 1u);
			idx = __shfl_sync(0xffu << (lane & 24), idx, lane & 24);
			if(idx >= a.npairs) done = true;
			else {
				pair = a.order[idx];
				qlen = a.qlen[pair]; tlen = a.tlen[pair];
				qs = a.seqs + a.qoff[pair]; ts = a.seqs + a.toff[pair];
				bw = a.bandwidth ? a.bandwidth : qlen;
				bw = (bw + kLanes - 1) / kLanes * kLanes;
				W = bw / kLanes;
				RS = bw * (PW + 1);
				tr = a.trace + a.trace_off[pair];
				meta = (int32_t*)(tr + (size_t)RS * (tlen + 1));
				row = 0; rbeg = 0; rslot = 0; mov = 0;
				best = kScoreMin; best_qe = 0; best_te = 0; stflag = 0;
				have = true;
				// ---- row -1 (bsalign.h:2094-2140), built directly in the circular row buffers --------------
				const bool two = (PW == 2);
				if(mode == 0 || mode == 2){
					int ext = two ? ge2 : ge1;
					int u0 = (int8_t)(go1 + ge1 + a.smin - a.smax);
					uint32_t xp = two ? (uint32_t)((go2 - go1) / (ge1 - ge2)) : 0;
					for(uint32_t p=t;p<bw;p+=kGroup){
						int v = ext;
						if(p == 0) v = u0; else if(two && p < xp) v = ge1;
						sU[p] = (int8_t)v;
					}
					// anchors: ub[0] = smax - smin, ub[j+1] = ub[j] + blocksum[j]
					for(int j=t;j<=kLanes;j+=kGroup){
						// sum of the first j*W cells
						int64_t n = (int64_t)j * W; // cells [0, n)
						int s = a.smax - a.smin;
						if(n > 0){
							s += u0;
							int64_t n1 = 0; // cells with ge1 among [1, n)
							if(two){ n1 = (int64_t)xp - 1; if(n1 > n - 1) n1 = n - 1; if(n1 < 0) n1 = 0; }
							s += (int)(n1 * ge1 + (n - 1 - n1) * ext);
						}
						sUB[j] = s;
					}
				} else {
					for(uint32_t p=t;p<bw;p+=kGroup) sU[p] = 0;
					for(int j=t;j<=kLanes;j+=kGroup) sUB[j] = 0;
				}
				if(PW >= 1) for(uint32_t p=t;p<bw;p+=kGroup) sE[p] = kEpi8Min;
				if(PW == 2) for(uint32_t p=t;p<bw;p+=kGroup) sQ[p] = kEpi8Min;
				for(uint32_t p=t;p<bw;p+=kGroup) sC[p] = p < qlen ? qs[p] : 4;
				if(t == 0){ sUB[17] = 0; sUB[18] = 0; sUB[19] = 0; }
			}
		}
		if(__all_sync(0xffffffffu, done)) break;
		__syncwarp();
		if(have && row == 0){
			// store row -1 to the trace (backcal may walk into it, bsalign.h:3922)
			for(uint32_t c=t;c<W;c+=kGroup){
				*(uint4*)(tr + 16 * c) = *(const uint4*)(sU + 16 * c);
				if(PW >= 1) *(uint4*)(tr + bw + 16 * c) = *(const uint4*)(sE + 16 * c);
				if(PW == 2) *(uint4*)(tr + 2 * bw + 16 * c) = *(const uint4*)(sQ + 16 * c);
			}
			if(t < 5) *(uint4*)(meta + 4 * t) = *(const uint4*)(sUB + 4 * t);
		}
		__syncwarp();

		// =============================== one DP row ================================================
		const uint32_t tb = have ? ts[row < tlen ? row : 0] : 0;
		int rh;
		if(mov && rbeg + bw < qlen){ // bsalign.h:3932-3946
			int lim = (int)qlen - (int)(rbeg + bw); if(lim < 0) lim = 0;
			if((uint32_t)lim < mov) mov = (uint32_t)lim;
		} else mov = 0;
		if(mov){
			rh = (mov - 1 < bw) ? group_getscore(sU, sUB, bw, W, rslot, mov - 1, t) : kScoreMin;
			if(mov - 1 >= bw) stflag |= 1;
		} else {
			if(rbeg) rh = kScoreMin;
			else if(mode == 1 || row == 0) rh = 0;
			else if(PW < 2) rh = (int)((uint32_t)go1 + (uint32_t)ge1 * row);
			else { uint32_t c1 = (uint32_t)go1 + (uint32_t)ge1 * row, c2 = (uint32_t)go2 + (uint32_t)ge2 * row; rh = (int)(c1 > c2 ? c1 : c2); }
		}
		// ---- band shift (bsalign.h:2244-2392) as index arithmetic -----------------------------------
		if(mov){
			if(mov >= bw){
				for(uint32_t p=t;p<bw;p+=kGroup){ sU[p] = 0; if(PW >= 1) sE[p] = 0; if(PW == 2) sQ[p] = 0; }
				for(int j=t;j<=kLanes;j+=kGroup) sUB[j] = kScoreMin;
				rbeg += mov; rslot = rbeg % bw;
				for(uint32_t p=t;p<bw;p+=kGroup){ uint32_t x = rbeg + p; uint32_t sl = rslot + p; if(sl >= bw) sl -= bw; sC[sl] = x < qlen ? qs[x] : 4; }
			} else {
				const uint32_t cyc = mov / W, mr = mov - cyc * W;
				// anchors of the old row advanced by the first mr cells of each block (:2310-2331)
				for(int j=t;j<kLanes;j+=kGroup){
					uint32_t s0 = rslot + j * W; if(s0 >= bw) s0 -= bw;
					int s = sUB[j];
					for(uint32_t p=0;p<mr;p++){ uint32_t sl = s0 + p; if(sl >= bw) sl -= bw; s += sU[sl]; }
					sTmp[j] = s;
				}
				const int ub16 = sUB[kLanes];
				__syncwarp(0xffu << (lane & 24));
				// overhang parameters (:2357-2369)
				uint32_t d; int c;
				if(PW == 2){ d = (uint32_t)((go1 - go2) / (ge2 - ge1)); c = min((int)a.smin, go2 + ge2) - 1 - a.smax + (go2 + ge2); }
				else { d = bw + 1; c = min((int)a.smin, go1 + ge1) - 1 - a.smax + (go1 + ge1); }
				const uint32_t i0 = bw - mov;
				for(int j=t;j<=kLanes;j+=kGroup){
					int v = (j + cyc < (uint32_t)kLanes) ? sTmp[j + cyc] : ub16;
					// block ends crossed by the overhang add the running overhang total (:2372-2389)
					uint32_t P = (uint32_t)j * W;
					if(j >= 1 && P > i0){
						uint32_t k = P - i0; // overhang cells in [i0, P)
						uint32_t n1 = (k - 1 < d - 1) ? k - 1 : d - 1;
						v += c + (int)n1 * ge1 + (int)(k - 1 - n1) * ge2;
					}
					sUB[j] = v;
				}
				// synthesized cells occupy the slots the dropped cells leave
				for(uint32_t k=t;k<mov;k+=kGroup){
					uint32_t sl = rslot + k; if(sl >= bw) sl -= bw;
					uint32_t x = rbeg + bw + k;
					sU[sl] = (int8_t)(k == 0 ? c : (k < d ? ge1 : ge2));
					if(PW >= 1) sE[sl] = 0;
					if(PW == 2) sQ[sl] = 0;
					sC[sl] = x < qlen ? qs[x] : 4;
				}
				rbeg += mov; rslot += mov; if(rslot >= bw) rslot -= bw;
			}
		}
		__syncwarp();

		// ---- cell 0 (bsalign.h:2899-2907) -----------------------------------------------------------
		const uint32_t T32 = colw[tb & 3];
		int h0;
		{
			int z0 = (int)(int8_t)__byte_perm(T32, 0xC1C1C1C1u, sC[rslot]);
			int u0 = sU[rslot], t0;
			h0 = (rh - sUB[0]) + z0;
			if(PW == 0) t0 = u0 + ge1;
			else if(PW == 1) t0 = u0 + sE[rslot];
			else t0 = u0 + max((int)sE[rslot], (int)sQ[rslot]);
			if(h0 >= t0){ if(h0 > kEpi8Max) h0 = kEpi8Max; } else h0 = kEpi8Min;
		}
		uint32_t sA0 = rslot + A * W; if(sA0 >= bw) sA0 -= bw;
		uint32_t sB0 = sA0 + W; if(sB0 >= bw) sB0 -= bw;

		// ---- pass 1: F (G) leaving every running block with nothing entering ------------------------
		uint32_t f = pk1(kEpi8Min), g = pk1(kEpi8Min);
		{
			uint32_t sA = sA0, sB = sB0;
			for(uint32_t i=0;i<W;i++){
				uint32_t u = pk_s8((uint8_t)sU[sA], (uint8_t)sU[sB]);
				uint32_t z = pk_s8(__byte_perm(T32, 0xC1C1C1C1u, sC[sA]), __byte_perm(T32, 0xC1C1C1C1u, sC[sB]));
				if(i == 0 && t == 0) z = (z & 0xffff0000u) | (uint32_t)(h0 & 0xffff);
				uint32_t h;
				if(PW == 0){
					uint32_t e = sadd(u, GE);
					h = smax3(e, z, f);
					f = ssub(sadd(h, GE), u);
				} else if(PW == 1){
					uint32_t e = sadd(pk_s8((uint8_t)sE[sA], (uint8_t)sE[sB]), u);
					h = smax3(e, z, f);
					f = ssub(smax(sadd(f, GE), sadd(h, GOE)), u);
				} else {
					uint32_t e = sadd(pk_s8((uint8_t)sE[sA], (uint8_t)sE[sB]), u);
					uint32_t q = sadd(pk_s8((uint8_t)sQ[sA], (uint8_t)sQ[sB]), u);
					h = smax(smax3(e, z, q), smax(f, g));
					uint32_t nu = __vneg2(u);
					uint32_t h2 = sadd(h, GOE);
					f = sadd(smax(sadd(f, GE), h2), nu);
					uint32_t h3 = sadd(h2, NGOQ);
					g = sadd(smax(sadd(g, GP), h3), nu);
				}
				if(++sA == bw) sA = 0;
				if(++sB == bw) sB = 0;
			}
		}
		sF[A] = (int8_t)lo16(f); sF[B] = (int8_t)hi16(f);
		if(PW == 2){ sF[16 + A] = (int8_t)lo16(g); sF[16 + B] = (int8_t)hi16(g); }
		__syncwarp();
		// ---- F penetration (bsalign.h:2639-2652): exact 16-step scalar scan, every thread redundantly ---
		{
			int finA = kEpi8Min, finB = kEpi8Min, ginA = kEpi8Min, ginB = kEpi8Min;
			int tW = (int)W * ge1, tW2 = (int)W * ge2;
			int ubp = sUB[0], ubn = sUB[1];
			int s = tW + kEpi8Min - (ubn - ubp), s2 = tW2 + kEpi8Min - (ubn - ubp);
			#pragma unroll
			for(int j=1;j<kLanes;j++){
				int fj = sF[j - 1];
				if(fj < s) fj = (int)(int8_t)s;
				int gj = 0;
				if(PW == 2){ gj = sF[16 + j - 1]; if(gj < s2) gj = (int)(int8_t)s2; }
				if(j == A){ finA = fj; ginA = gj; }
				if(j == B){ finB = fj; ginB = gj; }
				ubp = ubn; ubn = sUB[j + 1];
				s = tW + fj - (ubn - ubp);
				if(PW == 2) s2 = tW2 + gj - (ubn - ubp);
			}
			f = pk(finA, finB); g = pk(ginA, ginB);
		}
		// ---- pass 2: the row, written in place (bsalign.h:2934-2957 etc.) ------------------------------
		uint32_t v = 0, h = 0, u = 0, unew0 = 0;
		{
			uint32_t sA = sA0, sB = sB0;
			for(uint32_t i=0;i<W;i++){
				u = pk_s8((uint8_t)sU[sA], (uint8_t)sU[sB]);
				uint32_t z = pk_s8(__byte_perm(T32, 0xC1C1C1C1u, sC[sA]), __byte_perm(T32, 0xC1C1C1C1u, sC[sB]));
				if(i == 0 && t == 0) z = (z & 0xffff0000u) | (uint32_t)(h0 & 0xffff);
				uint32_t un, en = 0, qn = 0;
				if(PW == 0){
					uint32_t e = sadd(u, GE);
					h = smax3(e, z, f);
					un = ssub(h, v);
					v = ssub(h, u);
					f = ssub(sadd(h, GE), u);
				} else if(PW == 1){
					uint32_t e = sadd(pk_s8((uint8_t)sE[sA], (uint8_t)sE[sB]), u);
					h = smax3(e, z, f);
					un = ssub(h, v);
					uint32_t nu = __vneg2(u);
					v = sadd(h, nu);
					en = smax(ssub(sadd(e, GE), h), GOE);
					h = sadd(h, GOE);
					f = sadd(smax(sadd(f, GE), h), nu);
				} else {
					uint32_t e = sadd(pk_s8((uint8_t)sE[sA], (uint8_t)sE[sB]), u);
					uint32_t q = sadd(pk_s8((uint8_t)sQ[sA], (uint8_t)sQ[sB]), u);
					h = smax(smax3(e, z, q), smax(f, g));
					un = ssub(h, v);
					uint32_t nu = __vneg2(u), nh = __vneg2(h);
					v = sadd(h, nu);
					en = smax(sadd(sadd(e, GE), nh), GOE);
					qn = smax(sadd(sadd(q, GP), nh), GQP);
					h = sadd(h, GOE);
					f = sadd(smax(sadd(f, GE), h), nu);
					h = sadd(h, NGOQ);
					g = sadd(smax(sadd(g, GP), h), nu);
				}
				if(i == 0) unew0 = un;
				else { sU[sA] = (int8_t)lo16(un); sU[sB] = (int8_t)hi16(un); }
				if(PW >= 1){ sE[sA] = (int8_t)lo16(en); sE[sB] = (int8_t)hi16(en); }
				if(PW == 2){ sQ[sA] = (int8_t)lo16(qn); sQ[sB] = (int8_t)hi16(qn); }
				if(++sA == bw) sA = 0;
				if(++sB == bw) sB = 0;
			}
		}
		// ---- tail (bsalign.h:2618-2636) ------------------------------------------------------------------
		if(PW == 1) h = sadd(h, NGOE);
		else if(PW == 2) h = sadd(h, NGQP);
		const uint32_t vt = ssub(h, u);
		{
			int vtA = lo16(vt), vtB = hi16(vt);
			int vprev = __shfl_up_sync(0xffffffffu, vtB, 1, kGroup);
			if(t == 0) vprev = 0;
			int uA = clamp8(lo16(unew0) - vprev);
			int uB = clamp8(hi16(unew0) - vtA);
			__syncwarp();
			sUB[A + 1] += vtA;
			sUB[B + 1] += vtB;
			if(t == 0){ sUB[0] += uA; uA = 0; sUB[17] = (int32_t)rbeg; }
			sU[sA0] = (int8_t)uA;
			sU[sB0] = (int8_t)uB;
		}
		__syncwarp();
		// ---- stream the finished row to the traceback store ---------------------------------------------
		if(have){
			uint8_t *dst = tr + (size_t)RS * (row + 1);
			for(uint32_t c=t;c<W;c+=kGroup){
				*(uint4*)(dst + 16 * c) = *(const uint4*)(sU + 16 * c);
				if(PW >= 1) *(uint4*)(dst + bw + 16 * c) = *(const uint4*)(sE + 16 * c);
				if(PW == 2) *(uint4*)(dst + 2 * bw + 16 * c) = *(const uint4*)(sQ + 16 * c);
			}
			if(t < 5) *(uint4*)(meta + (size_t)kMetaInts * (row + 1) + 4 * t) = *(const uint4*)(sUB + 4 * t);
		}
		// ---- adaptive band steering (bsalign.h:3331-3349, 4005-4021) -------------------------------------
		{
			int rbx = 0;
			if(!(row <= W * kLanes / 4) && !(rbeg + W * kLanes >= qlen)){
				int noisy = 0, p0 = sUB[0];
				const int ub0 = p0;
				#pragma unroll
				for(int j=1;j<=kLanes;j++){ int p1 = sUB[j]; noisy += p1 < p0 ? p0 - p1 : p1 - p0; p0 = p1; }
				uint32_t nz = ((uint32_t)(noisy / kLanes)) / W * kLanes / 2;
				noisy = (int)(16u > nz ? 16u : nz);
				if(ub0 + noisy < p0) rbx = 2;
				else if(ub0 > p0 + noisy) rbx = 0;
				else rbx = 1;
			}
			if(mode == 0){
				int tq = (int)(tlen / qlen);
				int rbz = 2 * (tq > 1 ? tq : 1);
				int rby = (int)((1.0 * row / tlen) * qlen);
				if((int64_t)rbeg + rbz * (int64_t)(tlen - row - 1) + (int64_t)bw <= (int64_t)(uint32_t)(qlen + (uint32_t)rbz - 1)){
					uint32_t rem = tlen - row - 1;
					mov = 1 + ((qlen - (rbeg + bw)) / (rem > 1 ? rem : 1));
				} else if((int)rbeg < rby - (int)bw) mov = rbx + 1;
				else if((int)rbeg > rby) mov = rbx - 1 > 0 ? rbx - 1 : 0;
				else mov = rbx;
			} else mov = rbx;
		}
		// ---- end-point candidates (bsalign.h:4022-4045) ---------------------------------------------------
		if(mode != 0 && rbeg + bw >= qlen){
			int sc = group_getscore(sU, sUB, bw, W, rslot, qlen - 1 - rbeg, t);
			if(sc > best){ best = sc; best_qe = (int)qlen - 1; best_te = (int)row; }
		}
		row++;
		if(have && row == tlen){
			if(mode == 0){
				uint32_t pos = qlen - 1 - rbeg;
				if(pos < bw) best = group_getscore(sU, sUB, bw, W, rslot, pos, t);
				else { best = kScoreMin; stflag |= 1; }
				best_qe = (int)qlen - 1; best_te = (int)tlen - 1;
			} else {
				// row_max with the SSE reduction's tie-break order (bsalign.h:3213-3291)
				const uint32_t nchunk = (W + 31) / 32;
				#pragma unroll
				for(int which=0;which<2;which++){
					int j = which ? B : A;
					uint32_t s0 = which ? sB0 : sA0;
					int Max = kScoreMin, Scr = sUB[j]; uint32_t Idx = (uint32_t)j;
					uint32_t sl = s0;
					for(uint32_t c=0;c<nchunk;c++){
						uint32_t lo = c * 32, hi = lo + 32 < W ? lo + 32 : W;
						int run = 0, mx = -32767;
						for(uint32_t i=lo;i<hi;i++){ run += sU[sl]; if(run > mx) mx = run; if(++sl == bw) sl = 0; }
						int hh = Scr + mx;
						if(hh > Max){ Max = hh; Idx = (uint32_t)j | (c << 8); }
						Scr += run;
					}
					sRM[j] = Max; sRM[16 + j] = (int)Idx;
				}
				__syncwarp(0xffu << (lane & 24));
				int M4[4]; uint32_t I4[4];
				#pragma unroll
				for(int j=0;j<4;j++){
					int m0 = sRM[j], m1 = sRM[j + 8];
					uint32_t i0 = (uint32_t)sRM[16 + j], i1 = (uint32_t)sRM[16 + j + 8];
					if(sRM[j + 4] > m0){ m0 = sRM[j + 4]; i0 = (uint32_t)sRM[16 + j + 4]; }
					if(sRM[j + 12] > m1){ m1 = sRM[j + 12]; i1 = (uint32_t)sRM[16 + j + 12]; }
					if(m1 > m0){ m0 = m1; i0 = i1; }
					M4[j] = m0; I4[j] = i0;
				}
				int max_score = M4[0]; uint32_t bi = I4[0];
				#pragma unroll
				for(int j=1;j<4;j++) if(M4[j] > max_score){ max_score = M4[j]; bi = I4[j]; }
				if(max_score > best){
					uint32_t bl = bi & 0xff, bc = bi >> 8;
					uint32_t x = bc * 32, y = (bc + 1) * 32 < W ? (bc + 1) * 32 : W;
					uint32_t sl = rslot + bl * W + x; sl %= bw;
					uint32_t pos = x; int umax = kScoreMin, uscr = 0;
					for(;x<y;x++){ uscr += sU[sl]; if(uscr > umax){ pos = x; umax = uscr; } if(++sl == bw) sl = 0; }
					best = max_score; best_qe = (int)(rbeg + bl * W + pos); best_te = (int)tlen - 1;
				}
				__syncwarp(0xffu << (lane & 24));
			}
			if(t == 0){
				int32_t *rs = a.results + (size_t)pair * 10;
				rs[0] = best; rs[2] = best_qe; rs[4] = best_te;
				a.status[pair] = stflag;
			}
			have = false;
		}
		if(!have) mov = 0;
	}
}

// =====================================================================================================
// Traceback by re-derivation (bsalign.h:3667-3852), one thread per pair walking the HBM trace.
// =====================================================================================================
struct Epi8BtArgs {
	const uint8_t *seqs;
	const uint64_t *qoff, *toff;
	const uint32_t *qlen, *tlen;
	const uint32_t *order;
	uint32_t npairs;
	const uint8_t *trace;
	const uint64_t *trace_off;
	int32_t *results;
	int32_t *status;
	uint32_t *cigars;            // raw per-pair cigar regions (may be null)
	const uint64_t *cig_off;     // per pair offset into cigars (words); capacity = cig_off[i+1]-cig_off[i]
	uint32_t *dense;             // dense cigar arena: pairs append their final (reversed) cigar here
	uint64_t *dense_off;         // per pair: where its cigar starts in dense
	unsigned long long *dense_total;
	uint32_t *ncigar;
	uint32_t bandwidth;
	int mode;
	int pw;
	int8_t mtx[16];
	int8_t go1, ge1, go2, ge2;
};

struct TraceView {
	const uint8_t *tr; const int32_t *meta; uint32_t bw, W, RS; int tlen;
	__device__ __forceinline__ int beg(int row) const { return meta[(size_t)kMetaInts * (row + 1) + 17]; }
	__device__ __forceinline__ int ub(int row, int j) const { return meta[(size_t)kMetaInts * (row + 1) + j]; }
	__device__ __forceinline__ int cell(int row, int arr, uint32_t x) const { return (int)(int8_t)tr[(size_t)RS * (row + 1) + (size_t)arr * bw + (x % bw)]; }
	// H(col,row); sets err when the lookup leaves the stored band
	__device__ int score(int row, int col, int &err) const {
		if(row < -1 || row >= tlen){ err |= 1; return kScoreMin; }
		int b = beg(row);
		int64_t pos = (int64_t)col - b;
		if(pos < 0 || pos >= (int64_t)bw){ err |= 1; return kScoreMin; }
		uint32_t j = (uint32_t)pos / W;
		int s = ub(row, j);
		const uint8_t *r = tr + (size_t)RS * (row + 1);
		uint32_t sl = (uint32_t)(b + j * W) % bw;
		uint32_t n = (uint32_t)pos - j * W + 1;
		// head bytes up to 4-byte alignment, then whole words through dp4a
		while(n && (sl & 3)){ s += (int8_t)r[sl]; n--; if(++sl == bw) sl = 0; }
		while(n >= 4){
			s = __dp4a(*(const int*)(r + sl), 0x01010101, s);
			n -= 4; sl += 4; if(sl == bw) sl = 0;
		}
		while(n){ s += (int8_t)r[sl]; n--; if(++sl == bw) sl = 0; }
		return s;
	}
};

struct CigarSink {
	uint32_t *buf; uint32_t cap, n, run; int err;
	__device__ __forceinline__ void put(uint32_t w){ if(buf){ if(n < cap) buf[n] = w; else err |= 4; } n++; }
	__device__ __forceinline__ void push(uint32_t op, uint32_t sz){ // bsalign.h:409-417
		if(op == (run & 0xf)){ run += sz << 4; return; }
		if(run) put(run);
		run = sz << 4 | op;
	}
	__device__ __forceinline__ void flush(){ if(run) put(run); run = 0; }
};

// The walk emits operations end-to-start; the final cigar is their reverse (bsalign.h:3850, :1042).  Each pair
// reserves its exact length in one dense arena so that only real cigar words travel back over PCIe.
__device__ __forceinline__ void emit_dense(const CigarSink &cg, uint32_t *dense, uint64_t *dense_off, unsigned long long *dense_total, uint32_t *ncigar, uint32_t pair){
	if(ncigar) ncigar[pair] = cg.n;
	if(!cg.buf || !dense) return;
	uint32_t n = cg.n < cg.cap ? cg.n : cg.cap;
	unsigned long long off = atomicAdd(dense_total, (unsigned long long)n);
	dense_off[pair] = off;
	for(uint32_t i=0;i<n;i++) dense[off + i] = cg.buf[n - 1 - i];
}

__global__ void __launch_bounds__(128) epi8_backcal_kernel(const Epi8BtArgs a){
	uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if(idx >= a.npairs) return;
	const uint32_t pair = a.order[idx];
	const int qlen = (int)a.qlen[pair], tlen = (int)a.tlen[pair];
	const uint8_t *qs = a.seqs + a.qoff[pair], *ts = a.seqs + a.toff[pair];
	int32_t *rs = a.results + (size_t)pair * 10;
	const int mode = a.mode & 3, pw = a.pw;
	const int go1 = a.go1, ge1 = a.ge1, go2 = a.go2, ge2 = a.ge2;
	TraceView tv;
	tv.bw = a.bandwidth ? a.bandwidth : (uint32_t)qlen;
	tv.bw = (tv.bw + kLanes - 1) / kLanes * kLanes;
	tv.W = tv.bw / kLanes; tv.RS = tv.bw * (pw + 1); tv.tlen = tlen;
	tv.tr = a.trace + a.trace_off[pair];
	tv.meta = (const int32_t*)(tv.tr + (size_t)tv.RS * (tlen + 1));
	const int bw = (int)tv.bw;
	CigarSink cg;
	cg.buf = a.cigars ? a.cigars + a.cig_off[pair] : nullptr;
	cg.cap = a.cigars ? (uint32_t)(a.cig_off[pair + 1] - a.cig_off[pair]) : 0;
	cg.n = 0; cg.run = 0; cg.err = 0;
	int err = a.status[pair];
	int qb = rs[2], tb = rs[4];
	int mat = 0, mis = 0, ins = 0, del = 0, aln = 0;
	int Hcur, Hprev = 0, pend = 0, prior = 0;
	int64_t guard = 0; const int64_t guard_max = 8 * ((int64_t)qlen + tlen) + 64;
	const int qe = qb + 1, te = tb + 1;
	Hcur = tv.score(tb, qb, err);
	while(true){
		if(++guard > guard_max){ err |= 2; break; }
		if((pend & 0xf) == 2 || (pend & 0xf) == 4){
			int len = pend >> 4;
			Hprev = tv.score(tb, qb, err);
			int cost = ((pend & 0xf) == 2) ? go1 + len * ge1 : go2 + len * ge2;
			if(Hprev + cost == Hcur){
				cg.push(2, len);
				del += len; aln += len;
				Hcur = Hprev; pend = 0;
			} else { pend += 1 << 4; tb--; continue; }
		}
		if(qb < 0 || tb < 0) break;
		const int pbeg = tv.beg(tb - 1);
		if(qb == pbeg){
			if(qb){ Hprev = tv.ub(tb - 1, 0); prior = 0; }
			else if(mode == 1 || tb == 0) Hprev = 0;
			else if(pw < 2) Hprev = go1 + ge1 * tb;
			else Hprev = max(go1 + ge1 * tb, go2 + ge2 * tb);
		} else if(qb - pbeg <= bw){
			Hprev = tv.score(tb - 1, qb - 1, err);
		}
		{
			const int x = qb - pbeg;
			int bt, u = 0, e = 0, q = 0;
			if(x >= 0 && x < bw){
				u = tv.cell(tb - 1, 0, (uint32_t)qb);
				e = pw >= 1 ? tv.cell(tb - 1, 1, (uint32_t)qb) : (int)(int8_t)(go1 + ge1);
				q = pw == 2 ? tv.cell(tb - 1, 2, (uint32_t)qb) : 0;
			}
			const int s = a.mtx[qs[qb] * 4 + ts[tb]];
			const int h = Hcur - Hprev;
			if(x > bw) bt = 1;
			else if(x == bw) bt = (h == s) ? 0 : 1;
			else if(prior){
				if(h == s) bt = 0;
				else if(h == u + e) bt = 2;
				else if(pw == 2 && h == u + q) bt = 4;
				else bt = 1;
			} else {
				if(h == u + e) bt = 2;
				else if(pw == 2 && h == u + q) bt = 4;
				else if(h == s) bt = 0;
				else bt = 1;
			}
			prior = 1;
			if(bt == 0){
				if(qs[qb] == ts[tb]) mat++; else mis++;
				qb--; tb--; aln++;
				cg.push(0, 1);
				Hcur = Hprev;
			} else if(bt == 1){
				if(qb <= 0){
					cg.push(1, 1);
					Hcur = Hprev;
					qb--; ins++; aln++;
				} else {
					const int cbeg = tv.beg(tb);
					for(int sz=1;sz+cbeg<=qb;sz++){
						int tt = go1 + sz * ge1;
						if(pw == 2) tt = max(tt, go2 + sz * ge2);
						int Hl = tv.score(tb, qb - sz, err);
						if(Hl + tt == Hcur){
							cg.push(1, sz);
							Hcur = Hl; qb -= sz; ins += sz; aln += sz;
							break;
						}
					}
				}
			} else {
				pend = (1 << 4) | bt;
				tb--;
				continue;
			}
		}
	}
	if(mode == 1) cg.flush();
	else {
		uint32_t op = 0, sz = 0;
		if(qb >= 0){ op = 1; sz = qb + 1; ins += sz; qb = -1; }
		else if(tb >= 0){ op = 2; sz = tb + 1; del += sz; tb = -1; }
		aln += sz;
		cg.push(op, sz);
		cg.flush();
	}
	rs[1] = qb + 1; rs[2] = qe; rs[3] = tb + 1; rs[4] = te;
	rs[5] = mat; rs[6] = mis; rs[7] = ins; rs[8] = del; rs[9] = aln;
	emit_dense(cg, a.dense, a.dense_off, a.dense_total, a.ncigar, pair);
	a.status[pair] = err | cg.err;
}

} // namespace bsb200
