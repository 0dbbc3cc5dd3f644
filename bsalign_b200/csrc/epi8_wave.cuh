// epi8_wave.cuh -- single-pass "wavefront" forward kernel for full-band affine batches (replaces the two-pass row of
// bsalign.h:2885-2960 + the F hand-over of bsalign.h:2639-2652 where that is provably the same computation).
//
// The reference evaluates a row twice because its 16 SSE lanes run side by side: pass 1 finds the F leaving every lane's
// running block with nothing entering, a scalar scan hands F from block to block, pass 2 computes the row.  One step of the F
// chain is the map f -> min(127, max(f + (ge - u), beta)) (beta: the gap-open candidate of the cell, lower clamp included).
// Maps of that form are closed under composition, so a block is  F(f) = min(C, max(f + A, B))  with A = W*ge - sum(u) and, as
// long as the upper clamp never binds, the reference's hand-over  fin[j+1] = max(F(-63), fin[j] + A)  is just F(fin[j]) for
// fin[j] >= -63: the F that pass 2 itself leaves the block with.  So the lanes of a pair can run ONE pass each if lane j+1
// works one row behind lane j:
//   * thread t of the pair's group of 8 owns lanes 2t (low s16x2 half, row y - 2t) and 2t+1 (high half, row y - 2t - 1);
//     the block-exit F and the last cell's v travel to the next lane with a shuffle;
//   * SPLIT > 1 cuts every lane into SPLIT sub-blocks of whole 32-step groups, one more pipeline stage each (stage = SPLIT * lane +
//     sub-block, one row behind the stage before it): inside a lane that is nothing but the sequential chain of the SSE code cut
//     in pieces (F, the running -v and the running absolute score are handed on), and a pair is spread over 8 * SPLIT threads:
//     batches of a few long pairs (10 kb x 10 kb: HBM capacity seats ~5 pairs per SM) fill the SMs with 4x the warps;
//   * the two halves of a register sit on different target rows, so the substitution scores come from PRMT(column of row A,
//     column of row B, selector): scores are kept +63 (0..126), whose sign-replicating nibble yields the zero high byte and,
//     for positions past the query end, the -63 of the reference (0);
//   * in the trace the u and e pieces of a thread sit side by side (chunk c of a slot = 256 bytes: per thread 16 bytes of u, then its 16
//     bytes of e): the cell above a traceback step and the last chunk of its score lookup are then one 32-byte sector;
//   * the trace is written straight from the registers of the pass, in a SKEWED layout: slot s of a pair holds lane j's row
//     s - 1 - j, which is what the group produces together in one time step (one coalesced 128-byte line per chunk and
//     array); the traceback kernel adds the lane to the row index (TraceView::skew).
// Exactness is guarded, not assumed: a pair is flagged for the two-pass kernel (kStRedo) when (G1) an F of the pass reaches the
// upper clamp, (G2) a block-exit F drops under -63, or (G3) the u bytes of a finished lane do not add up to the difference of
// its anchors (the reference's scan uses the anchors, the pass the bytes: a saturated u or v makes them differ).  With G1-G3
// clean every value of the pass equals the reference's pass 2.  Full bands only (the band never moves: bsalign.h:3932 needs
// rbeg + bw < qlen), affine gaps with gape <= -1 and -64 <= gapo + gape <= -1, scores within +-63.
#pragma once
#include "epi8_forward.cuh"

namespace bsb200 {

struct WaveK { uint32_t C193, CM128, KGO, KUG, KHG, GOE129, GOE128, M1, ONE; };
struct WaveState { uint32_t f, h, u, nv; };

// ~x + g in both 16-bit halves as ONE IMAD (FMA pipe): x * 0xffffffff + K.  x * 0xffffffff + 0xffffffff is the bitwise complement;
// adding a negative g to its low half (0xffff - x, x small) always carries into the high half, g = 0 never does, so the high half's
// addend is pre-compensated.  g must be <= 0.
__host__ __device__ __forceinline__ uint32_t nadd_const(int g){
	const uint32_t lo = (uint32_t)g & 0xffffu, hi = (uint32_t)(g - (g < 0 ? 1 : 0)) & 0xffffu;
	return 0xffffffffu + ((hi << 16) | lo);
}

// One DP step of the thread's two lanes (bsalign.h:2934-2957 in difference form).  Values: u, e, f, h and the new u, e are kept + 128
// (unsigned bytes in the row images), the score + 63.  Against dp_step<1, true, true, true>:
//   * ev = e + u is a plain add (IMAD): its lower clamp at -128 cannot change anything because h >= z >= -63 and go >= -64;
//   * the gap-open candidate is not clamped (max(z, ev) >= -63, goe >= -64), which turns max(f + ge, max(z, ev) + goe) into
//     goe + max(f - go, max(z, ev)): one VIADDMNMX, the goe joins the subtraction of u in the IMAD that complements u;
//   * e' = max(ev + ge - h, goe): ge joins the complement of h in an IMAD, one VIADDMNMX.
// 8 ALU-pipe instructions (+ 4 IMAD) for two cells; the upper clamp of f and both clamps of u', v are the int8 saturation of the SSE code.
__device__ __forceinline__ void wave_step(WaveState &s, uint32_t u, uint32_t e, uint32_t z63, const WaveK &k, uint32_t &un, uint32_t &en){
	constexpr uint32_t C255 = 0x00ff00ffu, C128 = 0x00800080u;
	const uint32_t ev = u * k.ONE + e;                                      // e + u + 256
	const uint32_t hz = __viaddmax_s16x2(z63, k.C193, ev);                  // max(z, ev) + 256
	const uint32_t m = __viaddmax_s16x2(s.f, k.KGO, hz);                    // max(f - go, z, ev) + 256
	const uint32_t cug = u * k.M1 + k.KUG;                                  // ~u + goe + 1
	const uint32_t h = __viaddmax_s16x2(hz, k.CM128, s.f);                  // max(z, ev, f) + 128
	const uint32_t ch = not_fma(h, k.M1);
	const uint32_t chg = h * k.M1 + k.KHG;                                  // ~h + ge + 1
	un = __viaddmin_s16x2_relu(h, s.nv, C255);                              // subs(h, v) + 128
	s.nv = __viaddmin_s16x2(__viaddmax_s16x2(u, ch, kLO), kONE, C128);      // -subs(h, u)
	en = __viaddmax_s16x2(ev, chg, k.GOE128);                               // max(ev + ge - h, goe) + 128
	s.u = u; s.h = h;
	s.f = __viaddmin_s16x2_relu(m, cug, C255);                              // subs(max(f + ge, max(z, ev) + goe), u) + 128
}

// step K (0..7) of a 16-byte chunk in the step-major byte order, zero-extended: (lane A byte, lane B byte) -> s16x2
template<int K> __device__ __forceinline__ uint32_t entw(const uint4 &c){
	const uint32_t w = (K >> 1) == 0 ? c.x : (K >> 1) == 1 ? c.y : (K >> 1) == 2 ? c.z : c.w;
	return prmt(w, 0u, (K & 1) ? 0x4341u : 0x4240u);
}

// PRMT selector of one step: low byte fetches lane A's score from source a (the column of row A), high byte lane B's from
// source b; nibble 8|x replicates the sign of a byte < 128, i.e. yields 0: the high byte of each half, and the whole half for
// code 4 (past the query end: -63 + 63)
__device__ __forceinline__ uint32_t wsel(uint32_t cA, uint32_t cB){
	const uint32_t lo = cA < 4u ? (cA | ((8u | cA) << 4)) : 0x88u;
	const uint32_t hi = cB < 4u ? ((4u | cB) | ((12u | cB) << 4)) : 0xCCu;
	return lo | (hi << 8);
}

// VAR (experiments): bit 0: the selectors of odd steps are shifted down with IMAD.HI (FMA pipe) instead of SHF (ALU pipe)
// number of steps of a sub-block: the whole lane (in whole chunks) without a split, else whole 32-step groups
__host__ __device__ __forceinline__ uint32_t epi8_wave_block_steps(uint32_t W, uint32_t split){
	if(split <= 1) return (W + 7) / 8 * 8;
	return ((W + split - 1) / split + kStageAlign - 1) / kStageAlign * kStageAlign;
}
// a batch can be split when every sub-block of its narrowest band holds at least one step
__host__ __device__ __forceinline__ bool epi8_wave_split_ok(uint32_t minW, uint32_t split){ return split <= 1 || (split - 1) * epi8_wave_block_steps(minW, split) < minW; }
__host__ __device__ __forceinline__ uint32_t epi8_wave_slack(uint32_t split){ return kLanes * split - 1; }   // extra trace slots of a pair

template<bool ANCH, int SPLIT>
__global__ void __launch_bounds__(kFwdThreads, 4) epi8_wave_kernel(const Epi8Args a){
	extern __shared__ __align__(16) uint8_t smem_raw[];
	constexpr int kGT = 8 * SPLIT;                 // threads per pair
	const int lane = threadIdx.x & 31;
	const int g = lane & (kGT - 1), t = g & 7, b = g >> 3, gbase = lane - g;
	const unsigned gmask = (SPLIT == 4 ? 0xffffffffu : ((1u << kGT) - 1u)) << gbase;
	constexpr unsigned amask = 0xffffffffu;
	const int A = 2 * t, B = A + 1;
	const bool lane_end = (b == SPLIT - 1);
	const uint32_t IMG = a.max_img;
	uint8_t *gs = smem_raw + (size_t)(threadIdx.x / kGT) * a.group_smem;
	int8_t *sU = (int8_t*)gs;
	int8_t *sE = sU + IMG;
	uint8_t *sC = (uint8_t*)(sE + IMG);
	int32_t *sRM = (int32_t*)(sC + IMG);         // scratch words of the group: [0] global score, [1 + stage] maximum, [1 + 16 * SPLIT + stage] position

	const int mode = a.mode & 3;
	const int go1 = a.go1, ge1 = a.ge1;
	const int GOEi = (int8_t)(go1 + ge1);
	constexpr uint32_t C255 = 0x00ff00ffu, C129 = 0x00810081u;
	WaveK wk;
	wk.M1 = a.all_ones; wk.ONE = wk.M1 + 2u; wk.C193 = pk1(193); wk.CM128 = pk1(-128); wk.KGO = pk1(128 - go1); wk.KUG = nadd_const(GOEi + 1); wk.KHG = nadd_const(ge1 + 1);
	wk.GOE129 = pk1(GOEi + 129); wk.GOE128 = pk1(GOEi + 128);
	const uint32_t NGOE = pk1(-GOEi);
	const uint32_t K256 = a.c256;
	uint32_t colw[4];
	#pragma unroll
	for(int c=0;c<4;c++) colw[c] = (uint32_t)(uint8_t)(a.mtx[c] + 63) | ((uint32_t)(uint8_t)(a.mtx[4 + c] + 63) << 8) | ((uint32_t)(uint8_t)(a.mtx[8 + c] + 63) << 16) | ((uint32_t)(uint8_t)(a.mtx[12 + c] + 63) << 24);
	#define COLW(tb) (((tb) & 2) ? (((tb) & 1) ? colw[3] : colw[2]) : (((tb) & 1) ? colw[1] : colw[0]))
	// who hands over to this thread: inside a lane the thread of the previous sub-block (both halves), at a lane's start the last
	// sub-block of the previous lane - the other half of thread t's own lane pair for lane B, thread t-1's lane B for lane A
	const int srcA = b ? lane - 8 : (t ? gbase + 8 * (SPLIT - 1) + t - 1 : lane);
	const int srcB = b ? lane - 8 : gbase + 8 * (SPLIT - 1) + t;

	bool have = false, done = false;
	uint32_t pair = 0, qlen = 1, tlen = 1, W = 1, IB = 128, RS = 256;
	uint32_t s0 = 0, nst = 0, cb0 = 0, cb1 = 0;   // this thread's sub-block: first step, steps, chunk range
	const uint8_t *qs = a.seqs, *ts = a.seqs;
	uint8_t *tr = a.trace;
	int32_t *metaS = nullptr, *ub0p = nullptr;
	int T = 0;                               // time step of the pair: stage s works on row T - s
	int SA = 0, EA = 0, EB = 0;              // vertically accumulated anchors (bsalign.h:2618-2636): ub[0] (thread 0 of sub-block 0), ub[A+1], ub[B+1] (lane-end threads)
	uint32_t pubX1 = 0, pubY1 = 0; int pubX2 = 0, pubY2 = 0;   // what the stages after this thread's two take over: (F | v << 16, absolute score)
	uint32_t T32A = 0, T32B = 0, hist[SPLIT];
	int best = kScoreMin, best_te = 0, flagged = 0;
	uint32_t jq = 0, iq = 0;
	uint32_t n0 = 0; int c0u = 0, c0e = kEpi8Min;   // lane 0's first cell: selector nibble, u and e of the previous row (thread 0)
	int8_t *const rU = sU + 16 * t, *const rE = sE + 16 * t; uint8_t *const rC = sC + 16 * t;
	#define TOFF(i) ((((i) >> 3) << 7) + (((i) & 7) << 1))                       /* selectors: 2 bytes per step */
	#define WOFF(i) ((((i) >> 3) << 7) + ((((i) & 7) >> 1) << 2) + ((i) & 1))       /* u, e: lane A's byte of step i (lane B: + 2) */
	#define GOFF(i) ((((i) >> 3) << 8) + ((((i) & 7) >> 1) << 2) + ((i) & 1))       /* the same byte of u in the trace (e: + 16) */
	#define QCODE(x) ((x) < qlen ? (uint32_t)qs[(x)] : 4u)
	#pragma unroll
	for(int k=0;k<SPLIT;k++) hist[k] = 0;

	// A stage that has finished the pair's last row leaves its share of row_max (bsalign.h:3213-3263) in the group's scratch words
	// (its bytes in shared memory are overwritten by the idle time steps that follow): the maximum over its 32-step chunks - the
	// earliest chunk whose prefix maximum is strictly largest - and the first position inside that chunk that reaches it.
	auto stage_final = [&](int j, int anchor, int which){
		const int8_t *p = rU + 2 * which;
		int Max = kScoreMin, Scr = anchor; uint32_t bc = s0;
		for(uint32_t lo=s0;lo<s0+nst;lo+=32){
			const uint32_t hi = lo + 32 < s0 + nst ? lo + 32 : s0 + nst;
			int run = 0, mx = -32767;
			for(uint32_t i=lo;i<hi;i++){ run += (int)(uint8_t)p[WOFF(i)] - 128; if(run > mx) mx = run; }
			const int hh = Scr + mx;
			if(hh > Max){ Max = hh; bc = lo; }
			Scr += run;
		}
		uint32_t x = bc; const uint32_t y = bc + 32 < s0 + nst ? bc + 32 : s0 + nst;
		uint32_t pos = x; int umax = kScoreMin, uscr = 0;
		for(;x<y;x++){ uscr += (int)(uint8_t)p[WOFF(x)] - 128; if(uscr > umax){ pos = x; umax = uscr; } }
		sRM[1 + j * SPLIT + b] = Max; sRM[1 + kLanes * SPLIT + j * SPLIT + b] = (int)pos;
	};

	while(true){
		if(!have && !done){
			uint32_t idx = 0;
			if(g == 0) idx = atomicAdd(a.counter, 1u);
			idx = __shfl_sync(gmask, idx, gbase);
			if(idx >= a.npairs) done = true;
			else {
				pair = a.order[idx];
				qlen = a.qlen[pair]; tlen = a.tlen[pair];
				qs = a.seqs + a.qoff[pair]; ts = a.seqs + a.toff[pair];
				const uint32_t bw = ((a.bandwidth ? a.bandwidth : qlen) + kLanes - 1) / kLanes * kLanes;   // full band: no shorter than the query
				W = bw / kLanes;
				IB = epi8_image_bytes(W);
				RS = ANCH ? epi8_row_bytes(W, 1) : IB * 2;
				const uint32_t Wb = epi8_wave_block_steps(W, SPLIT);
				s0 = b * Wb; nst = s0 < W ? (W - s0 < Wb ? W - s0 : Wb) : 0;
				cb0 = s0 / 8; cb1 = (s0 + nst + 7) / 8;
				tr = a.trace + a.trace_off[pair];
				const uint32_t nslot = tlen + 1 + epi8_wave_slack(SPLIT);
				metaS = (int32_t*)(tr + (size_t)RS * nslot);
				ub0p = metaS + (size_t)16 * nslot;
				T = 0; have = true; flagged = a.force_redo ? 1 : 0;
				best = kScoreMin; best_te = 0;
				jq = (qlen - 1) / W; iq = (qlen - 1) - jq * W;
				pubX1 = pubY1 = (uint32_t)(kEpi8Min + 128); pubX2 = pubY2 = 0;
				T32A = COLW((uint32_t)ts[0]); T32B = T32A;   // stage 0's first row; the others re-load before they start
				// ---- row -1 of the thread's sub-blocks (bsalign.h:2094-2140) -----------------------------------
				const bool glob = (mode == 0 || mode == 2);
				const int u0 = (int8_t)(go1 + ge1 + a.smin - a.smax);
				// absolute score of row -1 at the end of band position p
				auto hinit = [&](int64_t p) -> int { return glob ? (int)(a.smax - a.smin + u0 + p * ge1) : 0; };
				for(uint32_t i=8*cb0;i<8*cb1;i++){
					uint32_t pA = A * W + i, pB = B * W + i;
					int vA = 0, vB = 0;
					if(glob){ vA = (pA == 0) ? u0 : ge1; vB = ge1; }
					if(i >= W){ vA = 0; vB = 0; }
					rU[WOFF(i)] = (int8_t)(vA + 128); rU[WOFF(i) + 2] = (int8_t)(vB + 128);
					rE[WOFF(i)] = (int8_t)(kEpi8Min + 128); rE[WOFF(i) + 2] = (int8_t)(kEpi8Min + 128);
					*(uint16_t*)(rC + TOFF(i)) = (uint16_t)(i < W ? wsel(QCODE(pA), QCODE(pB)) : wsel(4, 4));
				}
				SA = glob ? a.smax - a.smin : 0; EA = hinit((int64_t)(A + 1) * W - 1); EB = hinit((int64_t)(B + 1) * W - 1);
				n0 = (uint32_t)qs[0]; c0u = glob ? u0 : 0; c0e = kEpi8Min;
				// row -1 in the trace: a stage's row -1 is the slot of its number.  Lane A's pieces go out here (the B bytes of that slot
				// are never read); lane B's slot is written by the time step before lane B starts, which then puts its image back.
				{
					const uint32_t sA = SPLIT * A + b, sB = sA + SPLIT;
					uint8_t *d0 = tr + (size_t)RS * sA + 32 * t;
					for(uint32_t c=cb0;c<cb1;c++){
						*(uint4*)(d0 + 256 * c) = *(const uint4*)(rU + 128 * c);
						*(uint4*)(d0 + 256 * c + 16) = *(const uint4*)(rE + 128 * c);
					}
					if(lane_end){ metaS[(size_t)16 * sA + A] = EA; metaS[(size_t)16 * sB + B] = EB; }
					if(g == 0) ub0p[0] = SA;
					if(ANCH){
						for(uint32_t gi=s0/kAnchorSteps+1;gi*kAnchorSteps<=s0+nst&&gi<epi8_anchor_groups(W);gi++){
							*(int32_t*)(tr + (size_t)RS * sA + (size_t)IB * 2 + ((gi - 1) * 16 + A) * 4) = hinit((int64_t)A * W + gi * kAnchorSteps - 1);
							*(int32_t*)(tr + (size_t)RS * sB + (size_t)IB * 2 + ((gi - 1) * 16 + B) * 4) = hinit((int64_t)B * W + gi * kAnchorSteps - 1);
						}
					}
				}
			}
		}
		if(__all_sync(amask, done)) break;

		// =============================== one time step ================================================
		const int rA = T - (SPLIT * A + b), rB = rA - SPLIT;
		const bool actA = have && rA >= 0 && rA < (int)tlen, actB = have && rB >= 0 && rB < (int)tlen;
		// target base of lane A's next row, in flight during this step
		uint32_t tbn = 0;
		if(have && rA + 1 >= 0 && rA + 1 < (int)tlen) tbn = ts[rA + 1];
		// hand-over from the stages before this thread's two (they worked on the same rows one time step ago)
		const uint32_t inA1 = __shfl_sync(amask, pubX1, srcA), inB1 = __shfl_sync(amask, pubY1, srcB);
		const int inA2 = __shfl_sync(amask, pubX2, srcA), inB2 = __shfl_sync(amask, pubY2, srcB);
		if(actA || actB){
			// ---- cell 0 (bsalign.h:2899-2907): lane 0 only; the override of its score rides on the first chunk's first step ----
			uint32_t zm = 0xffffffffu, zo = 0u;
			if(g == 0){
				int rh;
				if(mode == 1 || rA == 0) rh = 0;
				else rh = (int)((uint32_t)go1 + (uint32_t)ge1 * (uint32_t)rA);
				const int z0 = (int)((T32A >> (8 * n0)) & 0xffu) - 63;
				const int t0 = c0u + c0e;
				int h0 = (rh - SA) + z0;
				if(h0 >= t0){ if(h0 > kEpi8Max) h0 = kEpi8Max; } else h0 = kEpi8Min;
				zm = 0xffff0000u; zo = (uint32_t)(h0 + 63);
			}
			WaveState st; st.h = 0; st.u = 0;
			uint32_t NV0 = 0u;
			if(b == 0){
				// start of a lane: F from the previous lane's exit (lane 0: -63), v starts at 0 and the first cell's u takes the previous
				// lane's last v afterwards: u = subs(u, v) (bsalign.h:2618-2636)
				const int finA = g ? (int)(inA1 & 0xffffu) : kEpi8Min + 128, vpA = g ? (int)(short)(inA1 >> 16) : 0;
				st.f = pk(finA, (int)(inB1 & 0xffffu)); st.nv = 0;
				NV0 = pk(-vpA, -(int)(short)(inB1 >> 16));
			} else {
				st.f = pk((int)(inA1 & 0xffffu), (int)(inB1 & 0xffffu));
				st.nv = pk((int)(short)(inA1 >> 16), (int)(short)(inB1 >> 16));
			}
			// absolute score at the start of the two sub-blocks (lane 0: the old ub[0]; its first u is re-based after the loop)
			const int ancA = g ? inA2 : SA, ancB = inB2;
			uint32_t gacc = st.f, fk = st.f, accA = 0, accB = 0;
			uint8_t *const gU = tr + (size_t)RS * (T + 1) + 32 * t;   // this thread's 32 bytes (u, e) of chunk 0 of the time step's slot
			#define WSTEP(K, LEFT) { if((K) < (LEFT)){ \
				/* (odd steps' selectors: SHF on the ALU pipe; IMAD.HI on the FMA pipe measured slower, c2 forward 53.9 -> 55.2 ms) */ \
				uint32_t z = prmt(T32A, T32B, ent_sel<K>(cs4)); \
				if((K) == 0) z = (z & zm) | zo; \
				wave_step(st, entw<K>(cu4), entw<K>(ce4), z, wk, un[K], en[K]); \
				if((K) & 1) gacc = __vimax3_s16x2(gacc, fk, st.f); else fk = st.f; } }
			#define WCHUNK(LEFT, RAGGED) { \
				const uint4 cu4 = *(const uint4*)(rU + 128 * c), cs4 = *(const uint4*)(rC + 128 * c), ce4 = *(const uint4*)(rE + 128 * c); \
				uint32_t un[8], en[8]; \
				if(RAGGED){ _Pragma("unroll") for(int k=0;k<8;k++){ un[k] = 0; en[k] = 0; } } \
				WSTEP(0, LEFT) WSTEP(1, LEFT) WSTEP(2, LEFT) WSTEP(3, LEFT) WSTEP(4, LEFT) WSTEP(5, LEFT) WSTEP(6, LEFT) WSTEP(7, LEFT) \
				un[0] = __viaddmin_s16x2_relu(un[0], NV0, C255); \
				zm = 0xffffffffu; zo = 0u; NV0 = 0u; \
				/* u + 128 and e + 128 lie in 0..255 in both halves: two steps interleave into a word with one IMAD (FMA pipe) */ \
				const uint4 ou4 = make_uint4(un[1] * K256 + un[0], un[3] * K256 + un[2], un[5] * K256 + un[4], un[7] * K256 + un[6]); \
				const uint4 oe4 = make_uint4(en[1] * K256 + en[0], en[3] * K256 + en[2], en[5] * K256 + en[4], en[7] * K256 + en[6]); \
				accA = __dp4a(ou4.x, 0x00000101u, accA); accB = __dp4a(ou4.x, 0x01010000u, accB); \
				accA = __dp4a(ou4.y, 0x00000101u, accA); accB = __dp4a(ou4.y, 0x01010000u, accB); \
				accA = __dp4a(ou4.z, 0x00000101u, accA); accB = __dp4a(ou4.z, 0x01010000u, accB); \
				accA = __dp4a(ou4.w, 0x00000101u, accA); accB = __dp4a(ou4.w, 0x01010000u, accB); \
				*(uint4*)(rU + 128 * c) = ou4; *(uint4*)(rE + 128 * c) = oe4; \
				*(uint4*)(gU + (size_t)c * 256) = ou4; *(uint4*)(gU + (size_t)c * 256 + 16) = oe4; \
				}
			const uint32_t nchunk = (W + 7) / 8, nfull = (s0 + nst) / 8;
			uint32_t c = cb0;
			if(ANCH){
				// the chunks of one anchor group at a time; the sub-lane anchor behind a group (H at the end of its last step = score at
				// the sub-block's start + its u bytes so far) is written between the groups, outside the chunk loop
				int32_t *an = (int32_t*)(tr + (size_t)RS * (T + 1) + (size_t)IB * 2) + (cb0 / kAnchorChunks) * 16;
				_Pragma("unroll 1")
				while(c < nfull){
					const uint32_t cend = c + kAnchorChunks < nfull ? c + kAnchorChunks : nfull;
					_Pragma("unroll 1")
					for(;c<cend;c++) WCHUNK(8u, false)
					if((c & (kAnchorChunks - 1)) == 0 && c < nchunk){
						const int corr = 128 * (int)(8 * (c - cb0));
						if(actA) an[A] = ancA + (int)accA - corr;
						if(actB) an[B] = ancB + (int)accB - corr;
						an += 16;
					}
				}
			} else {
				_Pragma("unroll 1")
				for(;c<nfull;c++) WCHUNK(8u, false)
			}
			if(c < cb1){ const uint32_t left = s0 + nst - 8 * c; WCHUNK(left, true) }
			#undef WCHUNK
			#undef WSTEP
			gacc = __vmaxs2(gacc, st.f);
			const int fxA = lo16(st.f), fxB = hi16(st.f);
			// absolute score at the end of the two sub-blocks
			int HA = ancA + (int)accA - 128 * (int)nst;
			const int HB = ancB + (int)accB - 128 * (int)nst;
			int xA = lo16(st.nv), xB = hi16(st.nv);      // what the next stage continues with: the running -v ...
			if(lane_end){
				// ... or, at the end of a lane, v of its last cell as the tail of the SSE code computes it (bsalign.h:2618-2636)
				const uint32_t yb = __viaddmax_s16x2(st.h, wk.GOE129, C129);
				uint32_t h = pk(lo16(yb) - 257, hi16(yb) - 257);
				const uint32_t ul = pk(lo16(st.u) - 128, hi16(st.u) - 128);
				h = sadd(h, NGOE);
				const uint32_t vt = ssubc(h, ~ul);
				xA = lo16(vt); xB = hi16(vt);
			}
			if(actB){
				if(fxB >= 255 || hi16(gacc) >= 255) flagged = 1;                                        // G1
				if(lane_end){
					EB += xB;
					if((t < 7 && fxB < kEpi8Min + 128) || HB != EB) flagged = 1;                         // G2, G3
					metaS[(size_t)16 * (T + 1) + B] = EB;
				}
				if(jq == (uint32_t)B && iq >= s0 && iq < s0 + nst && (mode != 0 || rB == (int)tlen - 1)){
					int sc = HB;
					for(uint32_t i=iq+1;i<s0+nst;i++) sc -= (int)(uint8_t)rU[WOFF(i) + 2] - 128;
					if(mode == 0) sRM[0] = sc;
					else if(sc > best){ best = sc; best_te = rB; }
				}
				if(mode != 0 && rB == (int)tlen - 1) stage_final(B, ancB, 1);
			}
			if(actA){
				if(g == 0){
					// lane 0 re-bases: ub[0] takes the first cell's u, which is stored as 0 (bsalign.h:2630-2634)
					const int dub0 = (int)(uint8_t)rU[0] - 128;
					rU[0] = (int8_t)128; gU[0] = 128;
					SA += dub0;
					ub0p[rA + 1] = SA;
					c0u = 0; c0e = (int)(uint8_t)rE[0] - 128;
				}
				if(fxA >= 255 || lo16(gacc) >= 255) flagged = 1;
				if(lane_end){
					EA += xA;
					if(fxA < kEpi8Min + 128 || HA != EA) flagged = 1;
					metaS[(size_t)16 * (T + 1) + A] = EA;
				}
				if(jq == (uint32_t)A && iq >= s0 && iq < s0 + nst && (mode != 0 || rA == (int)tlen - 1)){
					int sc = HA;
					for(uint32_t i=iq+1;i<s0+nst;i++) sc -= (int)(uint8_t)rU[WOFF(i)] - 128;
					if(mode == 0) sRM[0] = sc;
					else if(sc > best){ best = sc; best_te = rA; }
				}
				if(mode != 0 && rA == (int)tlen - 1) stage_final(A, g ? ancA : SA, 0);
			}
			// what the next stages take over: inside a lane the next sub-block of the same thread column, at a lane's end lane B goes to
			// thread t+1's lane A and lane A to this column's lane B
			const uint32_t pA1 = (uint32_t)(fxA & 0xffff) | ((uint32_t)xA << 16), pB1 = (uint32_t)(fxB & 0xffff) | ((uint32_t)xB << 16);
			const int pA2 = lane_end ? EA : HA, pB2 = lane_end ? EB : HB;
			if(lane_end){ pubX1 = pB1; pubX2 = pB2; pubY1 = pA1; pubY2 = pA2; }
			else { pubX1 = pA1; pubX2 = pA2; pubY1 = pB1; pubY2 = pB2; }
		}
		if(have && rB == -1){
			// lane B starts with the next time step: the steps before ran it on its row -1 image; put that image back (shared memory
			// and lane B's bytes of its row -1 slot, which is the slot of this time step)
			const bool glob = (mode == 0 || mode == 2);
			uint8_t *const gU = tr + (size_t)RS * (T + 1) + 32 * t;
			for(uint32_t i=8*cb0;i<8*cb1;i++){
				const int8_t ub_ = (int8_t)(((glob && i < W) ? ge1 : 0) + 128);
				rU[WOFF(i) + 2] = ub_; rE[WOFF(i) + 2] = (int8_t)(kEpi8Min + 128);
				gU[GOFF(i) + 2] = (uint8_t)ub_; gU[GOFF(i) + 18] = (uint8_t)(kEpi8Min + 128);
			}
		}
		#pragma unroll
		for(int k=SPLIT-1;k>0;k--) hist[k] = hist[k - 1];
		hist[0] = T32A;
		T32B = hist[SPLIT - 1]; T32A = COLW(tbn);
		T++;
		if(have && T == (int)tlen + kLanes * SPLIT - 1){
			// ---- the pair is through: every stage left its share of the last row in the group's scratch words ----
			flagged |= __shfl_xor_sync(gmask, flagged, 1);
			flagged |= __shfl_xor_sync(gmask, flagged, 2);
			flagged |= __shfl_xor_sync(gmask, flagged, 4);
			if(SPLIT >= 2) flagged |= __shfl_xor_sync(gmask, flagged, 8);
			if(SPLIT >= 4) flagged |= __shfl_xor_sync(gmask, flagged, 16);
			__syncwarp(gmask);
			int best_qe = (int)qlen - 1;
			if(mode == 0){
				best = sRM[0];
				best_te = (int)tlen - 1;
			} else {
				// the per-row candidates H(qlen-1, y) were collected by the thread that owns that stage
				const uint32_t Wb = epi8_wave_block_steps(W, SPLIT);
				const int own = gbase + 8 * (int)(iq / Wb) + (int)(jq >> 1);
				best = __shfl_sync(gmask, best, own);
				best_te = __shfl_sync(gmask, best_te, own);
				// row_max: per lane the earliest strictly largest chunk over its sub-blocks, then the lanes in the SSE code's tie-break
				// order (bsalign.h:3264-3291)
				auto lane_max = [&](int j, int &m, int &pos){
					m = sRM[1 + j * SPLIT]; pos = sRM[1 + kLanes * SPLIT + j * SPLIT];
					#pragma unroll
					for(int k=1;k<SPLIT;k++) if(sRM[1 + j * SPLIT + k] > m){ m = sRM[1 + j * SPLIT + k]; pos = sRM[1 + kLanes * SPLIT + j * SPLIT + k]; }
				};
				int M4[4], P4[4]; uint32_t I4[4];
				#pragma unroll
				for(int j=0;j<4;j++){
					int m0, p0, m1, p1, m2, p2, m3, p3;
					lane_max(j, m0, p0); lane_max(j + 4, m1, p1); lane_max(j + 8, m2, p2); lane_max(j + 12, m3, p3);
					uint32_t i0 = (uint32_t)j, i2 = (uint32_t)j + 8;
					if(m1 > m0){ m0 = m1; p0 = p1; i0 = (uint32_t)j + 4; }
					if(m3 > m2){ m2 = m3; p2 = p3; i2 = (uint32_t)j + 12; }
					if(m2 > m0){ m0 = m2; p0 = p2; i0 = i2; }
					M4[j] = m0; P4[j] = p0; I4[j] = i0;
				}
				int max_score = M4[0], bp = P4[0]; uint32_t bl = I4[0];
				#pragma unroll
				for(int j=1;j<4;j++) if(M4[j] > max_score){ max_score = M4[j]; bp = P4[j]; bl = I4[j]; }
				if(max_score > best){ best = max_score; best_qe = (int)(bl * W) + bp; best_te = (int)tlen - 1; }
			}
			if(g == 0){
				int32_t *rs = a.results + (size_t)pair * 10;
				rs[0] = best; rs[2] = best_qe; rs[4] = best_te;
				a.status[pair] = flagged ? kStRedo : kStSkew;
			}
			__syncwarp(gmask);
			have = false;
		}
	}
	#undef QCODE
	#undef TOFF
	#undef WOFF
	#undef GOFF
	#undef COLW
}

} // namespace bsb200
