// epi8_forward.cuh -- sm_100a forward kernel of the 8-bit banded DP (replaces bsalign.h:2094-3349, 3854-4045).
//
// Parallel decomposition:
//   * The reference's SSE word has 16 int8 lanes; lane j walks the "running block" of band positions
//     [j*W, (j+1)*W) sequentially, twice per row (pass 1: block-exit F; pass 2: the row).  To stay
//     bit-exact under int8 saturation that dependency structure is kept: one GROUP of 8 threads per pair,
//     each thread owning two SSE lanes packed as s16x2 in one register, so the native VIADDMNMX.S16x2 /
//     VIMNMX3.S16x2 instructions process two cells per issue with the saturation bounds of the SSE code.
//   * Row state lives in shared memory in a lane-pair layout: per array, chunk c holds for each thread t 16
//     bytes = the (lane 2t, lane 2t+1) byte pairs of steps 8c..8c+7, so the hot loops move 8 steps per 128-bit
//     LDS/STS and a group's access to a chunk touches every bank once (no conflicts).  The same image is what
//     is streamed to the HBM traceback store.
//   * The query profile of the reference (64 B per position) is replaced by a 2-byte PRMT selector per
//     step: one PRMT against the target base's matrix column yields both lanes' substitution scores.
//   * Four groups share a warp and run the row loop in lock step; groups fetch pairs from an atomic
//     counter (persistent scheduling, heaviest pairs first).
//   * FAST=true drops saturation bounds that provably cannot bind when all gap costs are <= 0 (the normal
//     case); FAST=false keeps every bound of the SSE code.  Both are bit-exact on their domain.
#pragma once
#include "common.cuh"

namespace bsb200 {

constexpr int kFwdThreads = 128;     // 16 groups per CTA

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
	uint32_t max_img;            // largest array image (bytes) in the batch (smem sizing)
	uint32_t group_smem;         // bytes of shared memory per group
	uint32_t gpw;                // groups per warp (4; fewer when a wide band makes shared memory the limit)
	int mode;
	int8_t mtx[16];
	int8_t go1, ge1, go2, ge2;
	int8_t smax, smin;
	uint32_t all_ones;           // 0xffffffff, passed at run time so that ~x can be issued as IMAD on the otherwise idle FMA pipe
	uint32_t c256, c65536;       // 256 and 65536 at run time: byte packing / half-word shifts as IMAD on the FMA pipe (wavefront kernel)
	int redo;                    // two-pass kernel: only take the pairs the wavefront kernel flagged (kStRedo)
	int force_redo;              // wavefront kernel: flag every pair (test hook: exercises the redo path)
	int bulk_store;              // two-pass kernel: finished row images leave shared memory as cp.async.bulk copies (BSB200_BULK_STORE)
};

// ---- bulk copy shared memory -> global memory (TMA unit, non-tensor form: SASS UBLKCP) ---------------------
// Both addresses 16-byte aligned, bytes a multiple of 16.  The writes that filled the shared-memory image went through the generic
// proxy: every writing thread issues bulk_fence() and the group synchronises before ONE thread issues the copy; that thread waits
// with bulk_wait_read() (source read, not global visibility: the kernel boundary orders the traceback behind it) before anyone
// overwrites the image.
__device__ __forceinline__ void bulk_fence(){ asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes){
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit(){ asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read(){ asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- saturating s16x2 arithmetic ------------------------------------------------------------------------
constexpr uint32_t kLO = 0xff80ff80u, kHI = 0x007f007fu, kONE = 0x00010001u;
__device__ __forceinline__ uint32_t sadd(uint32_t a, uint32_t b){ return __vmins2(__viaddmax_s16x2(a, b, kLO), kHI); }   // adds_epi8
__device__ __forceinline__ uint32_t sadd_lo(uint32_t a, uint32_t b){ return __viaddmax_s16x2(a, b, kLO); }             // upper bound cannot bind
// a - b with both bounds, given cb = ~b:  max(a + ~b, -129) + 1 = max(a - b, -128)
__device__ __forceinline__ uint32_t ssubc(uint32_t a, uint32_t cb){ return __viaddmin_s16x2(__viaddmax_s16x2(a, cb, 0xff7fff7fu), kONE, kHI); }  // subs_epi8
__device__ __forceinline__ uint32_t smax(uint32_t a, uint32_t b){ return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t smax3(uint32_t a, uint32_t b, uint32_t c){ return __vimax3_s16x2(a, b, c); }

// entry k (0..7) of a 16-byte chunk: (lane A byte, lane B byte) -> sign-extended s16x2
template<int K> __device__ __forceinline__ uint32_t ent(const uint4 &c){
	const uint32_t w = (K >> 1) == 0 ? c.x : (K >> 1) == 1 ? c.y : (K >> 1) == 2 ? c.z : c.w;
	return prmt(w, 0u, (K & 1) ? 0xB3A2u : 0x9180u);
}
// (w >> 16 as IMAD.HI on the FMA pipe instead of SHF was measured slower: 93.9 vs 90.8 ms on config 2)
template<int K> __device__ __forceinline__ uint32_t ent_sel(const uint4 &c){
	const uint32_t w = (K >> 1) == 0 ? c.x : (K >> 1) == 1 ? c.y : (K >> 1) == 2 ? c.z : c.w;
	return (K & 1) ? (w >> 16) : w;   // PRMT reads only bits 15:0 of its selector
}
// entry k as ZERO-extended halves (for the +128 biased u bytes of the FAST kernels)
template<int K> __device__ __forceinline__ uint32_t entz(const uint4 &c){
	const uint32_t w = (K >> 1) == 0 ? c.x : (K >> 1) == 1 ? c.y : (K >> 1) == 2 ? c.z : c.w;
	return prmt(w, 0u, (K & 1) ? 0x4342u : 0x4140u);
}
// low bytes of the two halves of ve (even entry) and vo (odd entry) -> one output word
__device__ __forceinline__ uint32_t pack2(uint32_t ve, uint32_t vo){ return prmt(ve, vo, 0x6420u); }

// PRMT selector that fetches (sext S[cA], sext S[cB]) from {column word, 0xC1C1C1C1}; code 4 = past the query end (-63)
__device__ __forceinline__ uint32_t zsel(uint32_t cA, uint32_t cB){ return cA | ((8u | cA) << 4) | (cB << 8) | ((8u | cB) << 12); }
// FAST kernels keep scores biased by +128 (unsigned): the selector fetches (S[cA]+128, 0, S[cB]+128, 0) from
// {column word + 128, 0x00000041}: byte 4 = 65 = -63 + 128 (past the query end), byte 5 = 0
__device__ __forceinline__ uint32_t zselb(uint32_t cA, uint32_t cB){ return cA | (5u << 4) | (cB << 8) | (5u << 12); }

struct RowState { uint32_t f, g, h, u, nv; };

// ~x = x * 0xffffffff + 0xffffffff (mod 2^32): an IMAD, i.e. FMA pipe, leaving the saturated ALU pipe to the DPX ops
__device__ __forceinline__ uint32_t not_fma(uint32_t x, uint32_t m1){ return x * m1 + m1; }

// one DP step for the thread's two lanes.  PASS2=false: only the F/G chain (pass 1).
// FAST=false: literal SSE arithmetic, every adds/subs with both saturation bounds, signed values.
// FAST=true (all gap costs <= 0): u, z, h, f, g and the new u are kept BIASED by +128, so that a two-sided saturating
// add of an unbiased term is ONE instruction (VIADDMNMX.RELU: min(a+b,255) then max(.,0)); bounds that cannot bind
// are dropped: e,q <= 0 so e+u never exceeds 127; x-h <= 0 because h >= e+u and ge <= 0; h+goe and f+ge cannot exceed 127.
// LAT=true (FAST, PW == 1; chosen for batches that leave an SM with a handful of warps, where the row loop is bound by
// the latency of the F chain and not by ALU throughput): the gap-open candidate is taken from max(ev, z) instead of
// max(ev, z, f).  Identical values (f + goe <= f + ge because go <= 0, and the lower clamp sits under both), but the
// loop-carried chain f -> h -> y -> f1 -> f (4 DPX ops) becomes f -> f1 -> f (2); pass 2 pays one more instruction.
// gap constants of dp_step as s16x2 (both halves equal).  The sums with 1 / 129 are built from the scalars here, once per kernel (or
// job): written as __vadd2(GOE, C129) inside dp_step, nvcc 12.9 split the constant into halves in one instantiation's ragged-chunk
// code and lost the +129 of the low half (caught by test_wide_bands / test_latency_and_throughput_variants).
struct DpK {
	uint32_t GE, GOE, GP, GQP, NGOQ, M1, Z0, GE1, GP1, GE129, GOE129, GP129;
	__device__ __forceinline__ void set(int ge1, int goe, int ge2, int gqp, int ngoq, uint32_t m1){
		GE = pk1(ge1); GOE = pk1(goe); GP = pk1(ge2); GQP = pk1(gqp); NGOQ = pk1(ngoq); M1 = m1;
		Z0 = m1 + 1u;   // 0 as a run-time register value: ptxas otherwise re-materialises the constant with a PRMT per use
		GE1 = pk1(ge1 + 1); GP1 = pk1(ge2 + 1); GE129 = pk1(ge1 + 129); GOE129 = pk1(goe + 129); GP129 = pk1(ge2 + 129);
	}
};

template<int PW, bool FAST, bool PASS2, bool LAT = false>
__device__ __forceinline__ void dp_step(RowState &s, uint32_t u, uint32_t e, uint32_t q, uint32_t z, const DpK &k, uint32_t &un, uint32_t &en, uint32_t &qn){
	const uint32_t GE = k.GE, GOE = k.GOE, GP = k.GP, GQP = k.GQP, NGOQ = k.NGOQ, M1 = k.M1, Z0 = k.Z0;
	if(FAST){
		constexpr uint32_t C129 = 0x00810081u, C255 = 0x00ff00ffu;
		// u, z, s.f, s.g biased; e, q, s.nv unbiased
		const uint32_t ev = __viaddmax_s16x2(u, PW == 0 ? GE : e, Z0);         // adds(e,u) + 128   (linear gaps: e = ge)
		uint32_t qv = 0, h;
		if(PW == 2){
			qv = __viaddmax_s16x2(u, q, Z0);
			h = smax(smax3(ev, z, qv), smax(s.f, s.g));
		} else h = smax3(ev, z, s.f);
		const uint32_t cu = not_fma(u, M1);
		if(LAT && PW == 1){
			const uint32_t hz = smax(ev, z);                                        // the part of h that does not hang on the F chain
			const uint32_t yz = __viaddmax_s16x2(hz, k.GOE129, C129);
			const uint32_t f1 = __viaddmax_s16x2(s.f, k.GE129, yz);
			if(PASS2){
				h = smax(hz, s.f);
				const uint32_t ch = not_fma(h, M1);
				un = __viaddmin_s16x2_relu(h, s.nv, C255);
				s.nv = __viaddmin_s16x2(__viaddmax_s16x2(u, ch, kLO), kONE, 0x00800080u);
				uint32_t x1 = __viaddmax_s16x2(ev, k.GE1, kONE);
				en = __viaddmax_s16x2(x1, ch, GOE);
				s.u = u;
				s.h = __viaddmax_s16x2(h, k.GOE129, C129);               // only the last step's value is read (row tail)
			}
			s.f = __viaddmin_s16x2_relu(f1, cu, C255);
			return;
		}
		if(PASS2){
			const uint32_t ch = not_fma(h, M1);
			un = __viaddmin_s16x2_relu(h, s.nv, C255);                              // subs(h, v) + 128
			s.nv = __viaddmin_s16x2(__viaddmax_s16x2(u, ch, kLO), kONE, 0x00800080u); // -(subs(h,u)): u_b + ~h_b = u - h - 1
			if(PW >= 1){
				uint32_t x1 = __viaddmax_s16x2(ev, k.GE1, kONE);        // adds(ev, ge) + 1 + 128
				en = __viaddmax_s16x2(x1, ch, GOE);                                 // max(x - h, goe): x1 + ~h_b = x - h
			}
			if(PW == 2){
				uint32_t x1 = __viaddmax_s16x2(qv, k.GP1, kONE);
				qn = __viaddmax_s16x2(x1, ch, GQP);
			}
			s.u = u;
		}
		if(PW == 0){
			s.h = h;
			uint32_t y = __viaddmax_s16x2(h, k.GE129, C129);            // adds(h, ge) + 128 + 129
			s.f = __viaddmin_s16x2_relu(y, cu, C255);                               // subs(y, u) + 128: y' + ~u_b = y - u + 128
		} else if(PW == 1){
			uint32_t y = __viaddmax_s16x2(h, k.GOE129, C129);           // adds(h, goe) + 128 + 129
			uint32_t f1 = __viaddmax_s16x2(s.f, k.GE129, y);            // max(adds(f, ge), y) + 128 + 129
			s.f = __viaddmin_s16x2_relu(f1, cu, C255);
			s.h = y;
		} else {
			uint32_t yb = __viaddmax_s16x2(h, GOE, Z0);                            // adds(h, goe) + 128
			uint32_t f1 = __viaddmax_s16x2(s.f, k.GE129, __vadd2(yb, C129));
			s.f = __viaddmin_s16x2_relu(f1, cu, C255);
			yb = __viaddmin_s16x2_relu(yb, NGOQ, C255);                             // subs(., goq) + 128
			uint32_t g1 = __viaddmax_s16x2(s.g, k.GP129, __vadd2(yb, C129));
			s.g = __viaddmin_s16x2_relu(g1, cu, C255);
			s.h = __vadd2(yb, C129);
		}
		return;
	}
	uint32_t ev, qv = 0, h;
	if(PW == 0) ev = sadd(u, GE);
	else ev = sadd(e, u);
	if(PW == 2){
		qv = sadd(q, u);
		h = smax(smax3(ev, z, qv), smax(s.f, s.g));
	} else h = smax3(ev, z, s.f);
	const uint32_t cu = ~u;
	if(PASS2){
		const uint32_t ch = ~h;
		un = sadd(h, s.nv);                                                    // u(x,y) = h - v(x-1,y)
		s.nv = __viaddmin_s16x2(__viaddmax_s16x2(u, ch, kLO), kONE, 0x00800080u);  // -(subs(h,u)) = clamp(u - h, -127, 128)
		if(PW >= 1) en = smax(ssubc(sadd(ev, GE), ch), GOE);
		if(PW == 2) qn = smax(ssubc(sadd(qv, GP), ch), GQP);
		s.u = u;
	}
	if(PW == 0){
		s.h = h;
		s.f = ssubc(sadd(h, GE), cu);
	} else {
		uint32_t y = sadd(h, GOE);
		s.f = ssubc(smax(sadd(s.f, GE), y), cu);
		if(PW == 2){
			y = sadd(y, NGOQ);
			s.g = ssubc(smax(sadd(s.g, GP), y), cu);
		}
		s.h = y; // the SSE code leaves h biased by the gap-open constant after the loop (:2958, :3177)
	}
}

// sum of entries [0, count) of lane j of a row image, spread over the group's threads
// (ubias = 128 when the image holds u + 128 as unsigned bytes)
__device__ __forceinline__ int group_lane_sum(const int8_t *img, uint32_t j, uint32_t count, int t, int ubias){
	const unsigned gm = 0xffu << ((threadIdx.x & 31) & 24);
	int s = 0;
	for(uint32_t k=t;k<count;k+=kGroup) s += (ubias ? (int)(uint8_t)img[epi8_cell_offset(j, k)] - ubias : (int)img[epi8_cell_offset(j, k)]);
	s += __shfl_xor_sync(gm, s, 1);
	s += __shfl_xor_sync(gm, s, 2);
	s += __shfl_xor_sync(gm, s, 4);
	return s;
}
// absolute H at band position pos (bsalign.h:3187-3197)
__device__ __forceinline__ int group_getscore(const int8_t *sU, const int32_t *sUB, uint32_t W, uint32_t pos, int t, int ubias){
	uint32_t j = pos / W, i = pos - j * W;
	return sUB[j] + group_lane_sum(sU, j, i + 1, t, ubias);
}

// ANCH: also write the sub-lane anchors (lanes longer than 64 steps; see common.cuh)
// NARROW: fewer than 4 groups per warp (very wide bands); the normal kernels keep compile-time full-warp masks
// LAT: latency-bound batches (a few warps per SM): short F chain (dp_step) and the next chunk's loads issued before the
// current chunk's arithmetic
// FULL: every pair's band covers its whole query (bandwidth 0, or a bandwidth no shorter than the longest query): the band never moves (bsalign.h:3932 needs rbeg + bw < qlen),
// so the shift and steering code is left out of the instantiation.  These kernels are compiled for 4 CTAs per SM (<= 128 registers;
// ptxas then takes ~115 instead of ~80 and schedules the hot loops better: c2 forward 82.5 -> 81.1 ms, 10 kb global 264 -> 248 ms);
// the same bound made the general kernels slower (c3 83.8 -> 90.6 ms), so they keep the default.
template<int PW, bool FAST, bool ANCH, bool NARROW, bool LAT = false, bool FULL = false>
__global__ void __launch_bounds__(kFwdThreads, FULL ? 4 : 0) epi8_forward_kernel(const Epi8Args a){
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int lane = threadIdx.x & 31;
	const int t = lane & 7;
	const unsigned gmask = 0xffu << (lane & 24);
	// wide bands: only the first gpw groups of a warp work (more warps then fit the shared memory); the others leave
	if(NARROW && (uint32_t)(lane >> 3) >= a.gpw) return;
	const unsigned amask = NARROW ? ((1u << (8 * a.gpw)) - 1u) : 0xffffffffu;
	const int A = 2 * t, B = A + 1;
	const uint32_t IMG = a.max_img;                 // bytes reserved per array image in shared memory
	uint8_t *gs = smem_raw + (size_t)(NARROW ? (threadIdx.x >> 5) * a.gpw + (lane >> 3) : (threadIdx.x >> 3)) * a.group_smem;
	int8_t *sU = (int8_t*)gs;
	int8_t *sE = sU + IMG;
	int8_t *sQ = sE + (PW >= 1 ? IMG : 0);
	uint8_t *sC = (uint8_t*)(sQ + (PW == 2 ? IMG : 0));   // PRMT selectors, 2 bytes per step
	int32_t *sUB = (int32_t*)(sC + IMG);            // kMetaInts ints: ub[17], rbeg
	int32_t *sTmp = sUB + kMetaInts;                // 20 ints scratch
	int8_t *sF = (int8_t*)(sTmp + kMetaInts);       // 16 fend, 16 gend
	int32_t *sRM = (int32_t*)(sF + 32);             // 32 ints scratch for row_max

	const int mode = a.mode & 3;
	const int go1 = a.go1, ge1 = a.ge1, go2 = a.go2, ge2 = a.ge2;
	const int GOEi = (int8_t)(go1 + ge1), GQPi = (int8_t)(go2 + ge2);
	const uint32_t GE = pk1(ge1), GOE = pk1(GOEi), GP = pk1(ge2), GQP = pk1(GQPi);
	const uint32_t M1 = a.all_ones;
	constexpr int UB = FAST ? 128 : 0;              // bias of the u bytes (and of z, h, f, g in registers)
	#define UBYTE(raw) (FAST ? (int)(uint8_t)(raw) - 128 : (int)(int8_t)(raw))
	#define ZSEL(ca, cb) (FAST ? zselb((ca), (cb)) : zsel((ca), (cb)))
	const uint32_t NGOE = pk1(-GOEi), NGOQ = pk1(-clamp8(GOEi - GQPi)), NGQP = pk1(-GQPi);
	DpK dpk; dpk.set(ge1, GOEi, ge2, GQPi, -clamp8(GOEi - GQPi), M1);
	// matrix columns: colw[tb] holds mtx[0*4+tb], mtx[1*4+tb], mtx[2*4+tb], mtx[3*4+tb] as bytes
	uint32_t colw[4];
	#pragma unroll
	for(int c=0;c<4;c++) colw[c] = (uint32_t)(uint8_t)(a.mtx[c] + UB) | ((uint32_t)(uint8_t)(a.mtx[4 + c] + UB) << 8) | ((uint32_t)(uint8_t)(a.mtx[8 + c] + UB) << 16) | ((uint32_t)(uint8_t)(a.mtx[12 + c] + UB) << 24);
	const uint32_t ZPAD = FAST ? 0x00000041u : 0xC1C1C1C1u;   // second PRMT source: the score past the query end (-63)

	bool have = false, done = false;
	uint32_t pair = 0, qlen = 1, tlen = 1, bw = 16, W = 1, IB = 128, row = 0, rbeg = 0, mov = 0;
	const uint8_t *qs = a.seqs, *ts = a.seqs;
	uint8_t *tr = a.trace;       // this pair's trace block (row -1 first)
	int32_t *meta = nullptr;     // this pair's anchors block
	uint32_t RS = 16;            // bytes per trace row
	uint32_t tb_next = 0;
	int best = kScoreMin, best_qe = 0, best_te = 0, stflag = 0;
	// this thread's 16 bytes inside chunk 0 of every array; chunk c is 128*c bytes further
	int8_t *const rU = sU + 16 * t, *const rE = sE + 16 * t, *const rQ = sQ + 16 * t; uint8_t *const rC = sC + 16 * t;
	// byte offset, relative to rU/rE/rQ/rC, of the thread's own step i (lane A; lane B is the next byte)
	#define TOFF(i) ((((i) >> 3) << 7) + (((i) & 7) << 1))

	// selector of query position x for this pair
	#define QCODE(x) ((x) < qlen ? (uint32_t)qs[(x)] : 4u)

	while(true){
		if(a.bulk_store){
			// the bulk copies of the previous row have read their images (they ran beside the steering and end-point code)
			if(t <= PW) bulk_wait_read();
			__syncwarp(gmask);
		}
		if(!have && !done){
			uint32_t idx = 0;
			if(t == 0) idx = atomicAdd(a.counter, 1u);
			idx = __shfl_sync(gmask, idx, lane & 24);
			if(idx >= a.npairs) done = true;
			else if(a.redo && !(a.status[a.order[idx]] & kStRedo)){ /* redo launch behind the wavefront kernel: this pair was not flagged */ }
			else {
				pair = a.order[idx];
				qlen = a.qlen[pair]; tlen = a.tlen[pair];
				qs = a.seqs + a.qoff[pair]; ts = a.seqs + a.toff[pair];
				bw = a.bandwidth ? a.bandwidth : qlen;
				bw = (bw + kLanes - 1) / kLanes * kLanes;
				W = bw / kLanes;
				IB = epi8_image_bytes(W);
				RS = ANCH ? epi8_row_bytes(W, PW) : IB * (PW + 1);
				tr = a.trace + a.trace_off[pair];
				meta = (int32_t*)(tr + (size_t)RS * (tlen + 1));
				row = 0; rbeg = 0; mov = 0;
				best = kScoreMin; best_qe = 0; best_te = 0; stflag = 0;
				have = true;
				tb_next = ts[0];
				// ---- row -1 (bsalign.h:2094-2140) ----------------------------------------------------------
				const bool two = (PW == 2);
				const bool glob = (mode == 0 || mode == 2);
				const int ext = two ? ge2 : ge1;
				const int u0 = (int8_t)(go1 + ge1 + a.smin - a.smax);
				const uint32_t xp = two ? (uint32_t)((go2 - go1) / (ge1 - ge2)) : 0;
				for(uint32_t i=0;i<IB/16;i++){
					uint32_t pA = A * W + i, pB = B * W + i;
					int vA = 0, vB = 0;
					if(glob){
						vA = (pA == 0) ? u0 : ((two && pA < xp) ? ge1 : ext);
						vB = (two && pB < xp) ? ge1 : ext;
					}
					if(i >= W){ vA = 0; vB = 0; }
					rU[TOFF(i)] = (int8_t)(vA + UB); rU[TOFF(i) + 1] = (int8_t)(vB + UB);
					if(PW >= 1){ rE[TOFF(i)] = kEpi8Min; rE[TOFF(i) + 1] = kEpi8Min; }
					if(PW == 2){ rQ[TOFF(i)] = kEpi8Min; rQ[TOFF(i) + 1] = kEpi8Min; }
					*(uint16_t*)(rC + TOFF(i)) = (uint16_t)(i < W ? ZSEL(QCODE(pA), QCODE(pB)) : ZSEL(4, 4));
				}
				for(int j=t;j<=kLanes;j+=kGroup){
					int s = 0;
					if(glob){
						int64_t n = (int64_t)j * W; // cells [0, n)
						s = a.smax - a.smin;
						if(n > 0){
							s += u0;
							int64_t n1 = 0; // cells holding ge1 among [1, n)
							if(two){ n1 = (int64_t)xp - 1; if(n1 > n - 1) n1 = n - 1; if(n1 < 0) n1 = 0; }
							s += (int)(n1 * ge1 + (n - 1 - n1) * ext);
						}
					}
					sUB[j] = s;
				}
				if(t == 0){ sUB[17] = 0; sUB[18] = 0; sUB[19] = 0; }
			}
		}
		if(__all_sync(amask, done)) break;
		__syncwarp(amask);
		if(have && row == 0){
			// store row -1 to the trace (backcal may walk into it, bsalign.h:3922)
			const uint32_t nch = IB / 16;
			for(uint32_t c=t;c<nch;c+=kGroup){
				*(uint4*)(tr + 16 * c) = *(const uint4*)(sU + 16 * c);
				if(PW >= 1) *(uint4*)(tr + IB + 16 * c) = *(const uint4*)(sE + 16 * c);
				if(PW == 2) *(uint4*)(tr + 2 * IB + 16 * c) = *(const uint4*)(sQ + 16 * c);
			}
			if(t < 5) *(uint4*)(meta + 4 * t) = *(const uint4*)(sUB + 4 * t);
			// sub-lane anchors of row -1
			for(uint32_t g=1;ANCH&&g<epi8_anchor_groups(W);g++){
				int sA_ = sUB[A], sB_ = sUB[B];
				for(uint32_t i=0;i<kAnchorSteps*g;i++){ sA_ += UBYTE(rU[TOFF(i)]); sB_ += UBYTE(rU[TOFF(i) + 1]); }
				*(int2*)(tr + (size_t)IB * (PW + 1) + ((g - 1) * 16 + A) * 4) = make_int2(sA_, sB_);
			}
		}
		__syncwarp(amask);

		// =============================== one DP row ================================================
		const uint32_t tb = tb_next;
		if(have && row + 1 < tlen) tb_next = ts[row + 1];
		int rh;
		if(!FULL && mov && rbeg + bw < qlen){ // bsalign.h:3932-3946
			int lim = (int)qlen - (int)(rbeg + bw); if(lim < 0) lim = 0;
			if((uint32_t)lim < mov) mov = (uint32_t)lim;
		} else mov = 0;
		if(!FULL && mov){
			if(mov <= 8 && mov <= W){
				// H at band position mov - 1, which lies in lane 0: every thread adds the same few cells (broadcast loads, no shuffles)
				rh = sUB[0];
				for(uint32_t k=0;k<mov;k++) rh += UBYTE(sU[2 * k]);
			} else rh = (mov - 1 < bw) ? group_getscore(sU, sUB, W, mov - 1, t, UB) : kScoreMin;
			if(mov - 1 >= bw) stflag |= 1;
		} else {
			if(rbeg) rh = kScoreMin;
			else if(mode == 1 || row == 0) rh = 0;
			else if(PW < 2) rh = (int)((uint32_t)go1 + (uint32_t)ge1 * row);
			else { uint32_t c1 = (uint32_t)go1 + (uint32_t)ge1 * row, c2 = (uint32_t)go2 + (uint32_t)ge2 * row; rh = (int)(c1 > c2 ? c1 : c2); }
		}
		// ---- band shift (bsalign.h:2244-2392) ----------------------------------------------------------
		if(!FULL && mov){
			if(mov >= bw){
				for(uint32_t i=0;i<IB/16;i++){
					uint32_t xA = rbeg + mov + A * W + i, xB = xA + W;
					*(uint16_t*)(rU + TOFF(i)) = (uint16_t)(UB | (UB << 8));
					if(PW >= 1) *(uint16_t*)(rE + TOFF(i)) = 0;
					if(PW == 2) *(uint16_t*)(rQ + TOFF(i)) = 0;
					*(uint16_t*)(rC + TOFF(i)) = (uint16_t)(i < W ? ZSEL(QCODE(xA), QCODE(xB)) : ZSEL(4, 4));
				}
				for(int j=t;j<=kLanes;j+=kGroup) sUB[j] = kScoreMin;
				rbeg += mov;
			} else {
				const uint32_t cyc = mov / W, mr = mov - cyc * W;
				// anchors of the old row advanced by the first mr cells of each block (:2310-2331)
				const int ub16 = sUB[kLanes];
				// overhang parameters (:2357-2369)
				uint32_t d; int c;
				if(PW == 2){ d = (uint32_t)((go1 - go2) / (ge2 - ge1)); c = min((int)a.smin, go2 + ge2) - 1 - a.smax + (go2 + ge2); }
				else { d = bw + 1; c = min((int)a.smin, go1 + ge1) - 1 - a.smax + (go1 + ge1); }
				const uint32_t i0 = bw - mov;   // first band position the old row does not cover
				if(cyc == 0 && mr <= 8){
					// fast path: the thread's byte-pair stream slides down by mr entries; the last mr entries come
					// from (own lane B, right neighbour's lane A) or, for the last lane, from the synthesized overhang
					uint4 fU = *(const uint4*)rU, fE = make_uint4(0, 0, 0, 0), fQ = fE, fC = *(const uint4*)rC;
					if(PW >= 1) fE = *(const uint4*)rE;
					if(PW == 2) fQ = *(const uint4*)rQ;
					// the neighbour's first mr entries: one word per array covers shifts of one or two cells
					uint4 nU = fU, nE = fE, nQ = fQ, nC = fC;
					nU.x = __shfl_down_sync(gmask, fU.x, 1, kGroup); nC.x = __shfl_down_sync(gmask, fC.x, 1, kGroup);
					if(PW >= 1) nE.x = __shfl_down_sync(gmask, fE.x, 1, kGroup);
					if(PW == 2) nQ.x = __shfl_down_sync(gmask, fQ.x, 1, kGroup);
					if(mr > 2){
						nU.y = __shfl_down_sync(gmask, fU.y, 1, kGroup); nU.z = __shfl_down_sync(gmask, fU.z, 1, kGroup); nU.w = __shfl_down_sync(gmask, fU.w, 1, kGroup);
						nC.y = __shfl_down_sync(gmask, fC.y, 1, kGroup); nC.z = __shfl_down_sync(gmask, fC.z, 1, kGroup); nC.w = __shfl_down_sync(gmask, fC.w, 1, kGroup);
						if(PW >= 1){ nE.y = __shfl_down_sync(gmask, fE.y, 1, kGroup); nE.z = __shfl_down_sync(gmask, fE.z, 1, kGroup); nE.w = __shfl_down_sync(gmask, fE.w, 1, kGroup); }
						if(PW == 2){ nQ.y = __shfl_down_sync(gmask, fQ.y, 1, kGroup); nQ.z = __shfl_down_sync(gmask, fQ.z, 1, kGroup); nQ.w = __shfl_down_sync(gmask, fQ.w, 1, kGroup); }
					}
					__syncwarp(gmask);
					// anchors (:2310-2331, :2372-2389): a lane's anchor advances by its first mr cells, which this thread holds in fU;
					// with a shift inside one lane only the end anchor ub[16] is reached by the overhang
					{
						int sumA = 0, sumB = 0;
						for(uint32_t k=0;k<mr;k++){
							const uint32_t w_ = (k >> 1) == 0 ? fU.x : (k >> 1) == 1 ? fU.y : (k >> 1) == 2 ? fU.z : fU.w, sh_ = (k & 1) * 16;
							sumA += UBYTE((w_ >> sh_) & 0xffu); sumB += UBYTE((w_ >> (sh_ + 8)) & 0xffu);
						}
						sUB[A] += sumA; sUB[B] += sumB;
						if(t == kGroup - 1){
							const uint32_t n1 = (mov - 1 < d - 1) ? mov - 1 : d - 1;
							sUB[kLanes] = ub16 + c + (int)n1 * ge1 + (int)(mov - 1 - n1) * ge2;
						}
					}
					// the thread's stream is 4 words per chunk; word w lives at chunk w/4, offset 4*(w%4)
					const uint32_t nW = IB / 32, ws = (2 * mr) / 4, bb = ((2 * mr) & 3) * 8;
					#define WOFF(w) ((((w) >> 2) << 7) + (((w) & 3) << 2))
					auto slide = [&](uint8_t *r){
						uint32_t lo = *(const uint32_t*)(r + WOFF(ws < nW ? ws : nW - 1));
						for(uint32_t w=0;w<nW;w++){
							uint32_t nx = w + ws + 1; if(nx >= nW) nx = nW - 1;
							uint32_t hi = *(const uint32_t*)(r + WOFF(nx));
							*(uint32_t*)(r + WOFF(w)) = __funnelshift_r(lo, hi, bb);
							lo = hi;
						}
					};
					#undef WOFF
					// shifts by one or two cells (what band_mov asks for; larger ones only come from the global-mode steering):
					// whole chunks, 128 bits per access, the shift as a clamped funnel over neighbouring words (16 or 32 bits)
					auto slide12 = [&](uint8_t *r){
						const uint32_t nC = IB / 128, sh = 16 * mr;
						uint4 cur = *(const uint4*)r;
						_Pragma("unroll 1")
						for(uint32_t c=0;c<nC;c++){
							uint4 nxt = cur;
							if(c + 1 < nC) nxt = *(const uint4*)(r + 128 * (c + 1));
							else nxt.x = cur.w;
							uint4 o;
							o.x = __funnelshift_rc(cur.x, cur.y, sh); o.y = __funnelshift_rc(cur.y, cur.z, sh);
							o.z = __funnelshift_rc(cur.z, cur.w, sh); o.w = __funnelshift_rc(cur.w, nxt.x, sh);
							*(uint4*)(r + 128 * c) = o;
							cur = nxt;
						}
					};
					if(mr <= 2){
						slide12((uint8_t*)rU); slide12((uint8_t*)rC);
						if(PW >= 1) slide12((uint8_t*)rE);
						if(PW == 2) slide12((uint8_t*)rQ);
					} else {
						slide((uint8_t*)rU); slide((uint8_t*)rC);
						if(PW >= 1) slide((uint8_t*)rE);
						if(PW == 2) slide((uint8_t*)rQ);
					}
					// entry k of a saved first chunk: byte pair (A, B)
					auto pairA = [](const uint4 &c4, uint32_t k){ uint32_t w = (k >> 1) == 0 ? c4.x : (k >> 1) == 1 ? c4.y : (k >> 1) == 2 ? c4.z : c4.w; return (w >> ((k & 1) * 16)) & 0xffu; };
					auto pairB = [](const uint4 &c4, uint32_t k){ uint32_t w = (k >> 1) == 0 ? c4.x : (k >> 1) == 1 ? c4.y : (k >> 1) == 2 ? c4.z : c4.w; return (w >> ((k & 1) * 16 + 8)) & 0xffu; };
					for(uint32_t k=0;k<mr;k++){
						const uint32_t i = W - mr + k;
						uint32_t uB, eB = 0, qB = 0, cB;
						if(t == 7){ // overhang cell k (:2370-2389)
							uB = (uint32_t)(uint8_t)((int8_t)(k == 0 ? c : (k < d ? ge1 : ge2)) + UB);
							cB = QCODE(rbeg + bw + k);
						} else {
							uB = pairA(nU, k); eB = pairA(nE, k); qB = pairA(nQ, k);
							uint32_t sel = (k & 1) ? (((k >> 1) == 0 ? nC.x : (k >> 1) == 1 ? nC.y : (k >> 1) == 2 ? nC.z : nC.w) >> 16) : ((k >> 1) == 0 ? nC.x : (k >> 1) == 1 ? nC.y : (k >> 1) == 2 ? nC.z : nC.w);
							cB = sel & 7u;          // lane A code of the neighbour's entry
						}
						uint32_t selo = (k & 1) ? (((k >> 1) == 0 ? fC.x : (k >> 1) == 1 ? fC.y : (k >> 1) == 2 ? fC.z : fC.w) >> 16) : ((k >> 1) == 0 ? fC.x : (k >> 1) == 1 ? fC.y : (k >> 1) == 2 ? fC.z : fC.w);
						uint32_t cA = (selo >> 8) & 7u; // own lane B code becomes lane A
						rU[TOFF(i)] = (int8_t)pairB(fU, k); rU[TOFF(i) + 1] = (int8_t)uB;
						if(PW >= 1){ rE[TOFF(i)] = (int8_t)pairB(fE, k); rE[TOFF(i) + 1] = (int8_t)eB; }
						if(PW == 2){ rQ[TOFF(i)] = (int8_t)pairB(fQ, k); rQ[TOFF(i) + 1] = (int8_t)qB; }
						*(uint16_t*)(rC + TOFF(i)) = (uint16_t)ZSEL(cA, cB);
					}
				} else {
					// general path (rare: global mode hurrying to the end): gather from the previous row's image in
					// the HBM trace, which this group finished writing in the previous iteration
					// anchors of the old row advanced by the first mr cells of each block (:2310-2331)
					for(int j=t;j<kLanes;j+=kGroup){
						int s = sUB[j];
						for(uint32_t k=0;k<mr;k++) s += UBYTE(sU[epi8_cell_offset(j, k)]);
						sTmp[j] = s;
					}
					__syncwarp(gmask);
					const uint8_t *pimg = tr + (size_t)RS * row; // image of row-1
					for(uint32_t i=0;i<W;i++){
						#pragma unroll
						for(int ln=0;ln<2;ln++){
							uint32_t P = (uint32_t)(A + ln) * W + i + mov;
							int uv, ev_ = 0, qv_ = 0;
							if(P < bw){
								uint32_t jo = P / W, io = P - jo * W;
								size_t off = epi8_cell_offset(jo, io);
								uv = (int8_t)pimg[off];
								if(PW >= 1) ev_ = (int8_t)pimg[IB + off];
								if(PW == 2) qv_ = (int8_t)pimg[2 * IB + off];
							} else {
								uint32_t k = P - bw;
								uv = (int8_t)(k == 0 ? c : (k < d ? ge1 : ge2)) + UB;
							}
							rU[TOFF(i) + ln] = (int8_t)uv;
							if(PW >= 1) rE[TOFF(i) + ln] = (int8_t)ev_;
							if(PW == 2) rQ[TOFF(i) + ln] = (int8_t)qv_;
						}
						uint32_t xA = rbeg + mov + A * W + i;
						*(uint16_t*)(rC + TOFF(i)) = (uint16_t)ZSEL(QCODE(xA), QCODE(xA + W));
					}
					__syncwarp(gmask);
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
				}
				rbeg += mov;
			}
		}
		__syncwarp(amask);

		// ---- cell 0 (bsalign.h:2899-2907) -----------------------------------------------------------
		// (FULL: selects over registers instead of the dynamic index, which puts colw in local memory and costs an LDL stall per row:
		// c2 forward 84.0 -> 82.5 ms.  The general instantiation keeps the indexed form: with the selects ptxas settles on 80 instead
		// of 94 registers and c3 runs 9 % slower, 83.7 -> 91.3 ms)
		const uint32_t T32 = FULL ? ((tb & 2) ? ((tb & 1) ? colw[3] : colw[2]) : ((tb & 1) ? colw[1] : colw[0])) : colw[tb & 3];
		int h0;
		{
			int z0 = UBYTE(prmt(T32, ZPAD, (uint32_t)sC[0] & 7u) & 0xffu);
			int u0 = UBYTE(sU[0]), t0;
			h0 = (rh - sUB[0]) + z0;
			if(PW == 0) t0 = u0 + ge1;
			else if(PW == 1) t0 = u0 + sE[0];
			else t0 = u0 + max((int)sE[0], (int)sQ[0]);
			if(h0 >= t0){ if(h0 > kEpi8Max) h0 = kEpi8Max; } else h0 = kEpi8Min;
		}
		const uint32_t zmask = t == 0 ? 0xffff0000u : 0xffffffffu, zor = t == 0 ? (uint32_t)((h0 + UB) & 0xffff) : 0u;
		const uint32_t nchunk = (W + 7) / 8, nfull = W / 8;

		// ---- pass 1: F (G) leaving every running block with nothing entering ------------------------
		// (full chunks run without per-step guards; a ragged last chunk, W % 8 != 0, takes the guarded copy)
		RowState st; st.f = pk1(kEpi8Min + UB); st.g = pk1(kEpi8Min + UB); st.h = 0; st.u = 0; st.nv = 0;
		{
			uint32_t dum0, dum1, dum2;
			#define P1STEP(K, LEFT) { if((K) < (LEFT)){ \
				uint32_t z = prmt(T32, ZPAD, ent_sel<K>(cs4)); \
				if((K) == 0 && c == 0) z = (z & zmask) | zor; \
				dp_step<PW, FAST, false, LAT>(st, FAST ? entz<K>(cu4) : ent<K>(cu4), ent<K>(ce4), ent<K>(cq4), z, dpk, dum0, dum1, dum2); } }
			#define P1BODY(LEFT) { P1STEP(0, LEFT) P1STEP(1, LEFT) P1STEP(2, LEFT) P1STEP(3, LEFT) P1STEP(4, LEFT) P1STEP(5, LEFT) P1STEP(6, LEFT) P1STEP(7, LEFT) }
			#define P1CHUNK(LEFT) { \
				const uint4 cu4 = *(const uint4*)(rU + 128 * c), cs4 = *(const uint4*)(rC + 128 * c); \
				uint4 ce4 = cu4, cq4 = cu4; \
				if(PW >= 1) ce4 = *(const uint4*)(rE + 128 * c); \
				if(PW == 2) cq4 = *(const uint4*)(rQ + 128 * c); \
				P1BODY(LEFT) }
			uint32_t c = 0;
			if(LAT){
				// chunk c + 1 is in flight while chunk c is computed; two register sets take turns (no copies)
				uint4 au, as_, ae, aq, bu, bs, be, bq;
				#define LDSET(U, S, E, Q, CI) { const uint32_t ci_ = (CI); U = *(const uint4*)(rU + 128 * ci_); S = *(const uint4*)(rC + 128 * ci_); \
					E = U; Q = U; if(PW >= 1) E = *(const uint4*)(rE + 128 * ci_); if(PW == 2) Q = *(const uint4*)(rQ + 128 * ci_); }
				#define P1ON(U, S, E, Q, LEFT) { const uint4 &cu4 = U, &cs4 = S, &ce4 = E, &cq4 = Q; P1BODY(LEFT) }
				LDSET(au, as_, ae, aq, 0u)
				_Pragma("unroll 1")
				while(c + 2 <= nfull){
					LDSET(bu, bs, be, bq, c + 1)
					P1ON(au, as_, ae, aq, 8u)
					c++;
					LDSET(au, as_, ae, aq, c + 1 < nchunk ? c + 1 : c)
					P1ON(bu, bs, be, bq, 8u)
					c++;
				}
				if(c < nfull){
					LDSET(bu, bs, be, bq, c + 1 < nchunk ? c + 1 : c)
					P1ON(au, as_, ae, aq, 8u)
					c++;
					if(c < nchunk){ const uint32_t left = W - 8 * c; P1ON(bu, bs, be, bq, left) }
				} else if(c < nchunk){ const uint32_t left = W - 8 * c; P1ON(au, as_, ae, aq, left) }
				#undef P1ON
			} else {
				_Pragma("unroll 1")   // one chunk per iteration keeps the loop body in the instruction cache (measured: -10% time)
				for(;c<nfull;c++) P1CHUNK(8u)
				if(c < nchunk){ const uint32_t left = W - 8 * c; P1CHUNK(left) }
			}
			#undef P1CHUNK
			#undef P1BODY
			#undef P1STEP
		}
		// ---- F penetration (bsalign.h:2639-2652) ---------------------------------------------------------
		// The reference hands F across the 16 blocks with a scalar scan: fin[0] = -63, fin[j+1] = max(fend[j], fin[j] + a[j]) with
		// a[j] = W*ge - (ub[j+1] - ub[j]), the sum truncated to int8 when it is taken.  Without the truncation this is a max-plus
		// recurrence, i.e. a prefix "product" of the maps x -> max(x + a, m): composed per thread (two lanes), scanned over the
		// group's 8 threads with shuffles (3 steps) instead of 15 dependent steps with shared-memory loads.  A taken sum always
		// exceeds a stored int8, so truncation can only bite when a sum exceeds 127: the group votes on that and only then runs
		// the literal scan.
		{
			const int ub0 = sUB[A], ub1 = sUB[B], ub2 = sUB[B + 1];
			constexpr int NEG = -(1 << 28);
			auto scan = [&](uint32_t fpk, int tW, int &finA, int &finB) -> bool {
				const int feA = lo16(fpk) - UB, feB = hi16(fpk) - UB;
				const int aA = tW - (ub1 - ub0), aB = tW - (ub2 - ub1);
				int GA = aA + aB, GM = max(feA + aB, feB);
				#pragma unroll
				for(int d=1;d<kGroup;d<<=1){
					const int pA = __shfl_up_sync(amask, GA, d, kGroup), pM = __shfl_up_sync(amask, GM, d, kGroup);
					if(t >= d){ GM = max(pM + GA, GM); GA += pA; }
				}
				int PA = __shfl_up_sync(amask, GA, 1, kGroup), PM = __shfl_up_sync(amask, GM, 1, kGroup);
				if(t == 0){ PA = 0; PM = NEG; }
				finA = max(kEpi8Min + PA, PM);
				const int sA = finA + aA;
				finB = max(sA, feA);
				const int sB = finB + aB;
				return max(sA, sB) > 127;
			};
			int finA, finB, ginA = kEpi8Min, ginB = kEpi8Min;
			bool ovf = scan(st.f, (int)W * ge1, finA, finB);
			if(PW == 2) ovf |= scan(st.g, (int)W * ge2, ginA, ginB);
			// (whole-warp shuffles and ballot: a partial mask costs a convergence check per call and runs the four groups one after the other)
			if((__ballot_sync(amask, ovf) >> (lane & 24)) & 0xffu){
				// literal 16-step scan with the int8 truncation, every thread redundantly
				sF[A] = (int8_t)(lo16(st.f) - UB); sF[B] = (int8_t)(hi16(st.f) - UB);
				if(PW == 2){ sF[16 + A] = (int8_t)(lo16(st.g) - UB); sF[16 + B] = (int8_t)(hi16(st.g) - UB); }
				__syncwarp(gmask);
				finA = finB = ginA = ginB = kEpi8Min;
				int tW = (int)W * ge1, tW2 = (int)W * ge2;
				int ubp = sUB[0], ubn = sUB[1];
				int s = tW + kEpi8Min - (ubn - ubp), s2 = tW2 + kEpi8Min - (ubn - ubp);
				#pragma unroll 1
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
				__syncwarp(gmask);
			}
			st.f = pk(finA + UB, finB + UB); st.g = pk(ginA + UB, ginB + UB);
		}
		// ---- pass 2: the row, written in place (bsalign.h:2934-2957 etc.) ------------------------------
		uint32_t unew0 = 0;
		st.nv = 0; st.h = 0; st.u = 0;
		{
			#define P2STEP(K, LEFT) { if((K) < (LEFT)){ \
				uint32_t z = prmt(T32, ZPAD, ent_sel<K>(cs4)); \
				if((K) == 0 && c == 0) z = (z & zmask) | zor; \
				dp_step<PW, FAST, true, LAT>(st, FAST ? entz<K>(cu4) : ent<K>(cu4), ent<K>(ce4), ent<K>(cq4), z, dpk, un[K], en[K], qn[K]); } }
			#define P2CHUNK(LEFT, RAGGED) { \
				const uint4 cu4 = *(const uint4*)(rU + 128 * c), cs4 = *(const uint4*)(rC + 128 * c); \
				uint4 ce4 = cu4, cq4 = cu4; \
				if(PW >= 1) ce4 = *(const uint4*)(rE + 128 * c); \
				if(PW == 2) cq4 = *(const uint4*)(rQ + 128 * c); \
				P2BODY(LEFT, RAGGED) }
			#define P2BODY(LEFT, RAGGED) { \
				uint32_t un[8], en[8], qn[8]; \
				if(RAGGED){ _Pragma("unroll") for(int k=0;k<8;k++){ un[k] = 0; en[k] = 0; qn[k] = 0; } } \
				P2STEP(0, LEFT) P2STEP(1, LEFT) P2STEP(2, LEFT) P2STEP(3, LEFT) P2STEP(4, LEFT) P2STEP(5, LEFT) P2STEP(6, LEFT) P2STEP(7, LEFT) \
				if(c == 0) unew0 = un[0]; \
				const uint4 ou4 = make_uint4(pack2(un[0], un[1]), pack2(un[2], un[3]), pack2(un[4], un[5]), pack2(un[6], un[7])); \
				*(uint4*)(rU + 128 * c) = ou4; \
				if(PW >= 1) *(uint4*)(rE + 128 * c) = make_uint4(pack2(en[0], en[1]), pack2(en[2], en[3]), pack2(en[4], en[5]), pack2(en[6], en[7])); \
				if(PW == 2) *(uint4*)(rQ + 128 * c) = make_uint4(pack2(qn[0], qn[1]), pack2(qn[2], qn[3]), pack2(qn[4], qn[5]), pack2(qn[6], qn[7])); }
			uint32_t c = 0;
			if(LAT){
				uint4 au, as_, ae, aq, bu, bs, be, bq;
				#define P2ON(U, S, E, Q, LEFT, RAGGED) { const uint4 &cu4 = U, &cs4 = S, &ce4 = E, &cq4 = Q; P2BODY(LEFT, RAGGED) }
				LDSET(au, as_, ae, aq, 0u)
				_Pragma("unroll 1")
				while(c + 2 <= nfull){
					LDSET(bu, bs, be, bq, c + 1)
					P2ON(au, as_, ae, aq, 8u, false)
					c++;
					LDSET(au, as_, ae, aq, c + 1 < nchunk ? c + 1 : c)   // (the row's last chunk re-reads itself before it is written)
					P2ON(bu, bs, be, bq, 8u, false)
					c++;
				}
				if(c < nfull){
					LDSET(bu, bs, be, bq, c + 1 < nchunk ? c + 1 : c)
					P2ON(au, as_, ae, aq, 8u, false)
					c++;
					if(c < nchunk){ const uint32_t left = W - 8 * c; P2ON(bu, bs, be, bq, left, true) }
				} else if(c < nchunk){ const uint32_t left = W - 8 * c; P2ON(au, as_, ae, aq, left, true) }
				#undef P2ON
				#undef LDSET
			} else {
				_Pragma("unroll 1")
				for(;c<nfull;c++) P2CHUNK(8u, false)
				if(c < nchunk){ const uint32_t left = W - 8 * c; P2CHUNK(left, true) }
			}
			#undef P2CHUNK
			#undef P2BODY
			#undef P2STEP
		}
		// ---- tail (bsalign.h:2618-2636) ------------------------------------------------------------------
		{
			// back to plain signed values (FAST keeps h biased by 128, or by 128+129 behind the gap-open add)
			constexpr int HB = FAST ? (PW == 0 ? 128 : 257) : 0;
			uint32_t h = pk(lo16(st.h) - HB, hi16(st.h) - HB);
			const uint32_t ul = pk(lo16(st.u) - UB, hi16(st.u) - UB);
			if(PW == 1) h = sadd(h, NGOE);
			else if(PW == 2) h = sadd(h, NGQP);
			const uint32_t vt = ssubc(h, ~ul);
			int vtA = lo16(vt), vtB = hi16(vt);
			int vprev = __shfl_up_sync(amask, vtB, 1, kGroup);
			if(t == 0) vprev = 0;
			int uA = clamp8(lo16(unew0) - UB - vprev);
			int uB = clamp8(hi16(unew0) - UB - vtA);
			__syncwarp(amask);
			sUB[A + 1] += vtA;
			sUB[B + 1] += vtB;
			if(t == 0){ sUB[0] += uA; uA = 0; sUB[17] = (int32_t)rbeg; }
			rU[0] = (int8_t)(uA + UB);
			rU[1] = (int8_t)(uB + UB);
		}
		__syncwarp(amask);
		if(ANCH && have){
			// sub-lane anchors (lanes longer than 32 steps): H at the end of step 32g-1 = lane anchor + the row's u bytes so far.
			// Done here, outside the hot loops (they are sensitive to code size), over the thread's own finished chunks;
			// IDP.4A issues on the FMA pipe.
			const uint32_t ng = epi8_anchor_groups(W);
			if(ng > 1){
				int32_t *an = (int32_t*)(tr + (size_t)RS * (row + 1) + (size_t)IB * (PW + 1));
				const int baseA = sUB[A], baseB = sUB[B];
				uint32_t accA = 0, accB = 0;
				_Pragma("unroll 1")
				for(uint32_t c=0;c<kAnchorChunks*(ng-1);c++){
					const uint4 w = *(const uint4*)(rU + 128 * c);
					if(FAST){
						accA = __dp4a(w.x, 0x00010001u, accA); accB = __dp4a(w.x, 0x01000100u, accB);
						accA = __dp4a(w.y, 0x00010001u, accA); accB = __dp4a(w.y, 0x01000100u, accB);
						accA = __dp4a(w.z, 0x00010001u, accA); accB = __dp4a(w.z, 0x01000100u, accB);
						accA = __dp4a(w.w, 0x00010001u, accA); accB = __dp4a(w.w, 0x01000100u, accB);
					} else {
						accA = (uint32_t)__dp4a((int)w.x, 0x00010001, (int)accA); accB = (uint32_t)__dp4a((int)w.x, 0x01000100, (int)accB);
						accA = (uint32_t)__dp4a((int)w.y, 0x00010001, (int)accA); accB = (uint32_t)__dp4a((int)w.y, 0x01000100, (int)accB);
						accA = (uint32_t)__dp4a((int)w.z, 0x00010001, (int)accA); accB = (uint32_t)__dp4a((int)w.z, 0x01000100, (int)accB);
						accA = (uint32_t)__dp4a((int)w.w, 0x00010001, (int)accA); accB = (uint32_t)__dp4a((int)w.w, 0x01000100, (int)accB);
					}
					if((c % kAnchorChunks) == kAnchorChunks - 1){
						const int corr = UB * (int)(8 * (c + 1));
						*(int2*)(an + (c / kAnchorChunks) * 16 + A) = make_int2(baseA + (int)accA - corr, baseB + (int)accB - corr);
					}
				}
			}
		}
		// ---- stream the finished row to the traceback store ---------------------------------------------
		if(have){
			uint8_t *dst = tr + (size_t)RS * (row + 1);
			const uint32_t nch = IB / 16; // 16-byte pieces per array image
			if(a.bulk_store){
				// one cp.async.bulk per array image (thread k: image k) instead of the LDS.128 -> STG.128 loop of the group
				bulk_fence();
				__syncwarp(gmask);
				if(t <= PW){ bulk_store_s2g(dst + (size_t)IB * t, sU + (size_t)IMG * t, IB); bulk_commit(); }
			} else
			for(uint32_t c=t;c<nch;c+=kGroup){
				*(uint4*)(dst + 16 * c) = *(const uint4*)(sU + 16 * c);
				if(PW >= 1) *(uint4*)(dst + IB + 16 * c) = *(const uint4*)(sE + 16 * c);
				if(PW == 2) *(uint4*)(dst + 2 * IB + 16 * c) = *(const uint4*)(sQ + 16 * c);
			}
			if(t < 5) *(uint4*)(meta + (size_t)kMetaInts * (row + 1) + 4 * t) = *(const uint4*)(sUB + 4 * t);
		}
		// ---- adaptive band steering (bsalign.h:3331-3349, 4005-4021) -------------------------------------
		if(!FULL){
			int rbx = 0;
			// sum of |ub[j] - ub[j-1]| (band_mov's noise term): each thread its two lanes, added up over the group with whole-warp
			// shuffles (hence outside the condition below) instead of 16 dependent shared-memory steps in every thread
			int noisy;
			{
				const int a_ = sUB[A], b_ = sUB[B], c_ = sUB[B + 1];
				noisy = abs(b_ - a_) + abs(c_ - b_);
				noisy += __shfl_xor_sync(amask, noisy, 1);
				noisy += __shfl_xor_sync(amask, noisy, 2);
				noisy += __shfl_xor_sync(amask, noisy, 4);
			}
			if(!(row <= W * kLanes / 4) && !(rbeg + W * kLanes >= qlen)){
				const int ub0 = sUB[0], p0 = sUB[kLanes];
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
			int sc = group_getscore(sU, sUB, W, qlen - 1 - rbeg, t, UB);
			if(sc > best){ best = sc; best_qe = (int)qlen - 1; best_te = (int)row; }
		}
		row++;
		if(have && row == tlen){
			if(mode == 0){
				uint32_t pos = qlen - 1 - rbeg;
				if(pos < bw) best = group_getscore(sU, sUB, W, pos, t, UB);
				else { best = kScoreMin; stflag |= 1; }
				best_qe = (int)qlen - 1; best_te = (int)tlen - 1;
			} else {
				// row_max with the SSE reduction's tie-break order (bsalign.h:3213-3291)
				const uint32_t nck = (W + 31) / 32;
				#pragma unroll
				for(int which=0;which<2;which++){
					int j = which ? B : A;
					const int8_t *p = rU + which;
					int Max = kScoreMin, Scr = sUB[j]; uint32_t Idx = (uint32_t)j;
					for(uint32_t c=0;c<nck;c++){
						uint32_t lo = c * 32, hi = lo + 32 < W ? lo + 32 : W;
						int run = 0, mx = -32767;
						for(uint32_t i=lo;i<hi;i++){ run += UBYTE(p[TOFF(i)]); if(run > mx) mx = run; }
						int hh = Scr + mx;
						if(hh > Max){ Max = hh; Idx = (uint32_t)j | (c << 8); }
						Scr += run;
					}
					sRM[j] = Max; sRM[16 + j] = (int)Idx;
				}
				__syncwarp(gmask);
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
					uint32_t pos = x; int umax = kScoreMin, uscr = 0;
					for(;x<y;x++){ uscr += UBYTE(sU[epi8_cell_offset(bl, x)]); if(uscr > umax){ pos = x; umax = uscr; } }
					best = max_score; best_qe = (int)(rbeg + bl * W + pos); best_te = (int)tlen - 1;
				}
				__syncwarp(gmask);
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
	#undef QCODE
	#undef TOFF
	#undef UBYTE
	#undef ZSEL
}

} // namespace bsb200
