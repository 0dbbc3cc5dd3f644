// poa_fast.cuh -- register-resident POA sweep kernel for the reference's default band (bandwidth 128, W = 8 steps per lane).
//
// Same contract as poa_sweep_kernel (poa_kernels.cuh; bspoa.h:2515-2618, 2232-2272; bsalign.h:2244-2392, 2474-2616, 2885-3179),
// for jobs whose band is exactly 128 cells, whose gap costs are all <= 0 with a gap-open cost (affine or two-piece) and whose
// bands never run past the read end (rpos + 128 <= slen for every node: the reference's clamp, bspoa.h:2170-2173).
//
//   * one group of 8 threads per job; thread t owns SSE lanes 2t and 2t+1 packed as s16x2, so the DPX instructions of
//     dp_step<PW, FAST> (epi8_forward.cuh) evaluate two cells per issue with the SSE saturation bounds;
//   * a row (u, e, q: 3 x 128 bytes) lives in REGISTERS: 4 words per array and thread = the (lane 2t, lane 2t+1) byte pairs
//     of the 8 steps, the same 16-byte image the pairwise forward kernel keeps per chunk; u is kept biased by +128;
//   * the band shift of row_movx is a byte funnel over the thread's 16 linear cells plus one SHFL from the right neighbour
//     (shifts of up to 4 cells, the common case: band offsets follow the consensus position); larger shifts go through
//     shared memory;
//   * the four query profiles are one byte per read position (PRMT selector nibbles) and two 4-byte score tables per edge;
//   * F penetration: the 16-step scalar scan of the reference only changes a lane's entry F when a gap crosses a whole lane;
//     the group first checks in parallel that it does not (two SHFLs + a vote) and only then falls back to the exact scan;
//   * the row just computed stays in registers and is the predecessor of the next node on the backbone, so HBM sees each
//     block written once (16-bit stores straight into the reference's striped layout) and read back only at graph joins.
#pragma once
#include "poa_kernels.cuh"
#include "epi8_forward.cuh"

namespace bsb200 {

constexpr int kPoaFastGroup = 8;
constexpr int kPoaFastBw = 128;
constexpr int kPoaFastSmem = 768;       // shared scratch bytes per group (slow paths only)

struct PoaFastArgs {
	PoaArgs base;
	const uint32_t *qsel;        // per read position: 0x80 | 17 * (base | hpc << 2), 16-byte aligned per job
	const uint64_t *qsel_off;    // per job byte offset into qsel
	uint32_t all_ones;           // 0xffffffff at run time (see Epi8Args::all_ones)
};

// one block per job: PRMT selector bytes of the read + traversal scratch
__global__ void poa_fast_prep_kernel(uint32_t njobs, const uint32_t *order, const uint8_t *queries, const uint64_t *qoff, const uint32_t *slen,
		uint8_t *qsel, const uint64_t *qsel_off, const uint64_t *node_off, int32_t *mpos, uint32_t *vst){
	for(uint32_t k=blockIdx.x;k<njobs;k+=gridDim.x){
		const uint32_t job = order[k];
		const uint64_t qo = qoff[job], so = qsel_off[job]; const uint32_t n = slen[job];
		for(uint32_t x=threadIdx.x;x<n+16;x+=blockDim.x){
			uint32_t c = 0;
			if(x < n){
				uint32_t b = queries[qo + x];
				c = (b & 3u) | ((x + 1 < n && queries[qo + x + 1] != b) ? 4u : 0u);   // set_query_prof_hpc, bsalign.h:2204-2206
			}
			qsel[so + x] = (uint8_t)(0x80u | (17u * c));
		}
		const uint64_t n0 = node_off[job], n1 = node_off[job + 1];
		for(uint64_t i=n0+threadIdx.x;i<n1;i+=blockDim.x){ mpos[i] = kPoaMposInit; vst[i] = 0; }
	}
}

struct Row8 {            // one 128-cell row in registers (interleaved byte pairs of lanes 2t / 2t+1, steps 0..7)
	uint32_t u0, u1, u2, u3;   // biased by +128
	uint32_t e0, e1, e2, e3;
	uint32_t q0, q1, q2, q3;
	int ubA, ubB, ub16;        // anchors of lanes 2t, 2t+1; ub16 is meaningful in thread 7 only
};

template<int PW>
__device__ __forceinline__ void row8_store(const Row8 &r, uint8_t *blk, int t){
	uint16_t *p = (uint16_t*)(blk + 2 * t);
	const uint32_t s0 = r.u0 ^ 0x80808080u, s1 = r.u1 ^ 0x80808080u, s2 = r.u2 ^ 0x80808080u, s3 = r.u3 ^ 0x80808080u;
	p[0] = (uint16_t)s0; p[8] = (uint16_t)(s0 >> 16); p[16] = (uint16_t)s1; p[24] = (uint16_t)(s1 >> 16);
	p[32] = (uint16_t)s2; p[40] = (uint16_t)(s2 >> 16); p[48] = (uint16_t)s3; p[56] = (uint16_t)(s3 >> 16);
	p += 64;
	p[0] = (uint16_t)r.e0; p[8] = (uint16_t)(r.e0 >> 16); p[16] = (uint16_t)r.e1; p[24] = (uint16_t)(r.e1 >> 16);
	p[32] = (uint16_t)r.e2; p[40] = (uint16_t)(r.e2 >> 16); p[48] = (uint16_t)r.e3; p[56] = (uint16_t)(r.e3 >> 16);
	if(PW == 2){
		p += 64;
		p[0] = (uint16_t)r.q0; p[8] = (uint16_t)(r.q0 >> 16); p[16] = (uint16_t)r.q1; p[24] = (uint16_t)(r.q1 >> 16);
		p[32] = (uint16_t)r.q2; p[40] = (uint16_t)(r.q2 >> 16); p[48] = (uint16_t)r.q3; p[56] = (uint16_t)(r.q3 >> 16);
	}
	int32_t *ub = (int32_t*)(blk + kPoaFastBw * (PW + 1));
	*(int2*)(ub + 2 * t) = make_int2(r.ubA, r.ubB);
	if(t == 7) ub[16] = r.ub16;
}

template<int PW>
__device__ __forceinline__ void row8_load(Row8 &r, const uint8_t *blk, int t){
	const uint16_t *p = (const uint16_t*)(blk + 2 * t);
	uint32_t a0 = p[0], a1 = p[8], a2 = p[16], a3 = p[24], a4 = p[32], a5 = p[40], a6 = p[48], a7 = p[56];
	const uint16_t *pe = p + 64;
	uint32_t b0 = pe[0], b1 = pe[8], b2 = pe[16], b3 = pe[24], b4 = pe[32], b5 = pe[40], b6 = pe[48], b7 = pe[56];
	uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
	if(PW == 2){ const uint16_t *pq = p + 128; c0 = pq[0]; c1 = pq[8]; c2 = pq[16]; c3 = pq[24]; c4 = pq[32]; c5 = pq[40]; c6 = pq[48]; c7 = pq[56]; }
	const int32_t *ub = (const int32_t*)(blk + kPoaFastBw * (PW + 1));
	const int2 ab = *(const int2*)(ub + 2 * t);
	r.ub16 = ub[16];
	r.ubA = ab.x; r.ubB = ab.y;
	r.u0 = (a0 | a1 << 16) ^ 0x80808080u; r.u1 = (a2 | a3 << 16) ^ 0x80808080u; r.u2 = (a4 | a5 << 16) ^ 0x80808080u; r.u3 = (a6 | a7 << 16) ^ 0x80808080u;
	r.e0 = b0 | b1 << 16; r.e1 = b2 | b3 << 16; r.e2 = b4 | b5 << 16; r.e3 = b6 | b7 << 16;
	r.q0 = c0 | c1 << 16; r.q1 = c2 | c3 << 16; r.q2 = c4 | c5 << 16; r.q3 = c6 | c7 << 16;
}

// interleaved pairs <-> the thread's 16 linear cells (A0..A7, B0..B7)
__device__ __forceinline__ void deinterleave(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t &l0, uint32_t &l1, uint32_t &l2, uint32_t &l3){
	l0 = prmt(w0, w1, 0x6420u); l1 = prmt(w2, w3, 0x6420u); l2 = prmt(w0, w1, 0x7531u); l3 = prmt(w2, w3, 0x7531u);
}
__device__ __forceinline__ void interleave(uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3){
	w0 = prmt(l0, l2, 0x5140u); w1 = prmt(l0, l2, 0x7362u); w2 = prmt(l1, l3, 0x5140u); w3 = prmt(l1, l3, 0x7362u);
}
// shift the thread's linear cells down by r bytes (1..4); fill = the 4 cells that follow them (right neighbour / overhang)
__device__ __forceinline__ void slide4(uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3, uint32_t r8, unsigned gmask, int t, uint32_t over){
	uint32_t l0, l1, l2, l3;
	deinterleave(w0, w1, w2, w3, l0, l1, l2, l3);
	uint32_t l4 = __shfl_down_sync(gmask, l0, 1, kPoaFastGroup);
	if(t == kPoaFastGroup - 1) l4 = over;
	const uint32_t n0 = __funnelshift_rc(l0, l1, r8), n1 = __funnelshift_rc(l1, l2, r8), n2 = __funnelshift_rc(l2, l3, r8), n3 = __funnelshift_rc(l3, l4, r8);
	interleave(n0, n1, n2, n3, w0, w1, w2, w3);
}

// GPW = jobs (groups of 8 threads) per warp.  The sweep of one job is a sequential chain of row operations; a warp that carries one
// job issues every instruction for 8 of its 32 lanes.  With GPW > 1 the groups of a warp run a flattened state machine: phase A (per
// group, divergent) finds the group's next graph edge -- popping nodes, handling edges into the tail, finishing and fetching jobs --
// and phase B (all groups converged) performs one row update (+ merge) per group, so the hot code is issued once for up to 4 jobs.
template<int PW, int GPW>
__global__ void __launch_bounds__(32) poa_sweep_fast_kernel(const PoaFastArgs fa){
	extern __shared__ __align__(16) uint8_t poa_fsm[];
	const PoaArgs &a = fa.base;
	const int lane = threadIdx.x & 31;
	if(lane >= kPoaFastGroup * GPW) return;
	const int t = lane & 7;
	const unsigned gmask = 0xffu << (lane & 24);
	const unsigned amask = GPW == 4 ? 0xffffffffu : ((1u << (kPoaFastGroup * GPW)) - 1u);
	const int A = 2 * t, B = A + 1;
	// shared scratch of the slow paths, per group: 3 x 128 linear cells, 17 anchors, 32 F bytes, 32 ints for row_max, 16 deltas
	uint8_t *gsm = poa_fsm + (size_t)(lane >> 3) * kPoaFastSmem;
	int8_t *sLin = (int8_t*)gsm;
	int32_t *sUB = (int32_t*)(gsm + 384);
	int8_t *sF = (int8_t*)(sUB + 20);
	int32_t *sRM = (int32_t*)(sF + 32);
	int32_t *sD = sRM + 32;                                 // 16 anchor deltas for the exact F scan
	const uint32_t M1 = fa.all_ones;
	const uint32_t Z0 = M1 + 1u;                            // 0 as a run-time register value (see dp_step)
	constexpr uint32_t bw = kPoaFastBw, W = 8;
	constexpr uint32_t mmblk = (bw * (PW + 1) + 68 + 15) / 16 * 16;

	// ---- per-group job state -------------------------------------------------------------------------------------------
	bool have = false, done = false;
	uint32_t job = 0;
	int mode = 1, Mm = 0, Xx = 0, O = 0, E = 0, Q = 0, P = 0, T = 0, refbonus = 0, go1 = 0, ge1 = 0, go2 = 0, ge2 = 0;
	uint32_t GE = 0, GOE = 0, GP = 0, GQP = 0, NGOE = 0, NGOQ = 0, NGQP = 0;
	DpK dpk; dpk.set(0, 0, 0, 0, 0, M1);
	int smax_nt = 0, smin_nt = 0;
	uint32_t slen = 0;
	const uint32_t *qsel = fa.qsel;
	const int2 *node = a.node; const int32_t *eoff = a.eoff, *edst = a.edst;
	int32_t *mpos = a.mpos; uint32_t *vst = a.vst, *stack = a.stack;
	uint32_t head = 0, tail = 0;
	uint8_t *rows = a.rows;
	int maxscr = kScoreMin, maxidx = -1, maxoff = -1, stflag = 0;
	unsigned long long nupd = 0, nmrg = 0;
	uint32_t ovd = bw + 1, over_u = 0; int ovc = 0;
	#define POA_OVER(k) ((int)(int8_t)((k) == 0 ? ovc : ((uint32_t)(k) < ovd ? ge1 : ge2)))
	Row8 Pd, Cu;
	Pd.u0 = Pd.u1 = Pd.u2 = Pd.u3 = Pd.e0 = Pd.e1 = Pd.e2 = Pd.e3 = Pd.q0 = Pd.q1 = Pd.q2 = Pd.q3 = 0u; Pd.ubA = Pd.ubB = Pd.ub16 = 0;
	Cu = Pd;
	uint32_t cur_node = 0xffffffffu;
	// traversal cursor: node being expanded and its remaining out-edges; the top of the stack is kept in a register
	uint32_t sp = 0, tos = 0, u = 0;
	int ei = 0, e1 = 0, rpos_u = 0, mpos_u = 0;
	uint32_t base_u = 0;

	// H at band position pos of a row: anchor + the lane's first cells, computed by the owner and broadcast
	auto getscore = [&](const Row8 &r, int64_t pos) -> int {
		if(pos < 0 || pos >= (int64_t)bw){ stflag |= 1; return kScoreMin; }
		const uint32_t jj = (uint32_t)pos >> 3, ii = (uint32_t)pos & 7u;
		uint32_t l0, l1, l2, l3;
		deinterleave(r.u0, r.u1, r.u2, r.u3, l0, l1, l2, l3);
		const uint32_t lo = (jj & 1) ? l2 : l0, hi = (jj & 1) ? l3 : l1;
		const uint32_t n = ii + 1;                         // cells to sum
		const uint32_t wlo = n >= 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - n)));
		const uint32_t whi = n <= 4 ? 0u : (0x01010101u >> (8 * (8 - n)));
		int s = ((jj & 1) ? r.ubB : r.ubA) + (int)__dp4a(lo, wlo, 0u) + (int)__dp4a(hi, whi, 0u) - 128 * (int)n;
		return __shfl_sync(gmask, s, (int)(jj >> 1), kPoaFastGroup);
	};


	for(;;){
		// =========================== phase A: this group's next edge into a non-tail node =============================
		uint32_t v = 0; int mpos_v = 0;
		bool act = false;
		while(!done){
			if(!have){
				uint32_t idx = 0;
				if(t == 0) idx = atomicAdd(a.counter, 1u);
				idx = __shfl_sync(gmask, idx, 0, kPoaFastGroup);
				if(idx >= a.njobs){ done = true; break; }
				job = a.order[idx];
				const int32_t *par = a.par + (size_t)job * 10;
				mode = par[1] & 3; Mm = par[2]; Xx = par[3]; O = par[4]; E = par[5]; Q = par[6]; P = par[7]; T = par[8]; refbonus = par[9];
				go1 = (int8_t)O; ge1 = (int8_t)E; go2 = (int8_t)Q; ge2 = (int8_t)P;
				const int GOEi = (int8_t)(go1 + ge1), GQPi = (int8_t)(go2 + ge2);
				GE = pk1(ge1); GOE = pk1(GOEi); GP = pk1(ge2); GQP = pk1(GQPi);
				NGOE = pk1(-GOEi); NGOQ = pk1(-clamp8(GOEi - GQPi)); NGQP = pk1(-GQPi);
				dpk.set(ge1, GOEi, ge2, GQPi, -clamp8(GOEi - GQPi), M1);
				smax_nt = (int8_t)(Mm + refbonus + 1); smin_nt = (int8_t)Xx;
				slen = a.slen[job];
				qsel = (const uint32_t*)((const uint8_t*)fa.qsel + fa.qsel_off[job]);
				const uint64_t n0 = a.node_off[job];
				node = a.node + n0; eoff = a.eoff + n0 + job; edst = a.edst + a.edge_off[job];
				mpos = a.mpos + n0; vst = a.vst + n0; stack = a.stack + n0;
				head = a.head[job]; tail = a.tail[job];
				rows = a.rows + a.row_off[job];
				maxscr = kScoreMin; maxidx = -1; maxoff = -1; stflag = 0; nupd = 0; nmrg = 0;
				if(PW == 2){ ovd = (uint32_t)((go1 - go2) / (ge2 - ge1)); ovc = min(smin_nt, go2 + ge2) - 1 - smax_nt + (go2 + ge2); }
				else { ovd = bw + 1; ovc = min(smin_nt, go1 + ge1) - 1 - smax_nt + (go1 + ge1); }
				// the 4 overhang cells behind the band, biased (fill of the last thread's slide)
				over_u = (uint32_t)((POA_OVER(0) + 128) & 0xff) | (uint32_t)((POA_OVER(1) + 128) & 0xff) << 8 |
					(uint32_t)((POA_OVER(2) + 128) & 0xff) << 16 | (uint32_t)((POA_OVER(3) + 128) & 0xff) << 24;
				// ---- head row (bsalign.h:2094-2140) ----------------------------------------------------------------
				{
					const bool two = (PW == 2), glob = (mode == 0 || mode == 2);
					const int ext = two ? ge2 : ge1;
					const int u0 = (int8_t)(go1 + ge1 + smin_nt - smax_nt);
					const uint32_t xp = two ? (uint32_t)((go2 - go1) / (ge1 - ge2)) : 0;
					uint32_t w[4];
					#pragma unroll
					for(int k=0;k<4;k++){
						uint32_t word = 0;
						#pragma unroll
						for(int h=0;h<4;h++){
							const uint32_t i = 2 * k + (h >> 1), p = (uint32_t)((h & 1) ? B : A) * W + i;
							int v = 0;
							if(glob) v = (p == 0) ? u0 : ((two && p < xp) ? ge1 : ext);
							word |= (uint32_t)((v + 128) & 0xff) << (8 * h);
						}
						w[k] = word;
					}
					Cu.u0 = w[0]; Cu.u1 = w[1]; Cu.u2 = w[2]; Cu.u3 = w[3];
					const uint32_t mn = 0x01010101u * (uint32_t)(uint8_t)kEpi8Min;
					Cu.e0 = Cu.e1 = Cu.e2 = Cu.e3 = mn;
					Cu.q0 = Cu.q1 = Cu.q2 = Cu.q3 = (PW == 2) ? mn : 0u;
					auto anchor = [&](int k) -> int {
						int s = 0;
						if(glob){
							int64_t n = (int64_t)k * W;
							s = smax_nt - smin_nt;
							if(n > 0){
								s += u0;
								int64_t n1 = 0;
								if(two){ n1 = (int64_t)xp - 1; if(n1 > n - 1) n1 = n - 1; if(n1 < 0) n1 = 0; }
								s += (int)(n1 * ge1 + (n - 1 - n1) * ext);
							}
						}
						return s;
					};
					Cu.ubA = anchor(A); Cu.ubB = anchor(B); Cu.ub16 = anchor(16);
					row8_store<PW>(Cu, rows + (size_t)head * mmblk, t);
					cur_node = head;
				}

				if(mode != 1) mpos[head] = -1;
				sp = 1; tos = head;
				ei = 0; e1 = 0;
				have = true;
			}
			if(ei < e1){
				v = (uint32_t)edst[ei++];
				if(mode != 1){
					mpos_v = mpos[v];
					if(mpos_u + 1 < mpos_v){ mpos_v = mpos_u + 1; mpos[v] = mpos_v; }
				}
				if(v != tail){ act = true; break; }
				// end candidates (bspoa.h:2548-2580)
				int moff = (int)min((int64_t)slen, (int64_t)rpos_u + (int64_t)bw) - 1;
				int smx = getscore(Pd, (int64_t)moff - rpos_u);
				if((int)slen > moff + 1){
					uint32_t rem = slen - (uint32_t)moff - 1u;
					if(PW < 2) smx += (int)((uint32_t)O + (uint32_t)E * rem);
					else { uint32_t c1 = (uint32_t)O + (uint32_t)E * rem, c2 = (uint32_t)Q + (uint32_t)P * rem; smx += (int)(c1 > c2 ? c1 : c2); }
				}
				smx += T;
				if(smx > maxscr){ maxscr = smx; maxidx = (int)u; maxoff = moff; }
				if(mode == 1){
					// row_max with the SSE reduction's tie-break order (bsalign.h:3213-3291); rare (edges into the tail): via shared memory
					uint32_t l0, l1, l2, l3;
					deinterleave(Pd.u0 ^ 0x80808080u, Pd.u1 ^ 0x80808080u, Pd.u2 ^ 0x80808080u, Pd.u3 ^ 0x80808080u, l0, l1, l2, l3);
					__syncwarp(gmask);
					*(uint4*)(sLin + 16 * t) = make_uint4(l0, l1, l2, l3);
					sUB[A] = Pd.ubA; sUB[B] = Pd.ubB;
					__syncwarp(gmask);
					#pragma unroll
					for(int which=0;which<2;which++){
						const int j = which ? B : A;
						int run = 0, mx = -32767;
						for(uint32_t i=0;i<W;i++){ run += sLin[j * 8 + i]; if(run > mx) mx = run; }
						sRM[j] = max(kScoreMin, sUB[j] + mx); sRM[16 + j] = j;      // a single chunk of <= 32 steps
					}
					__syncwarp(gmask);
					int M4[4]; uint32_t I4[4];
					#pragma unroll
					for(int k=0;k<4;k++){
						int m0 = sRM[k], m1 = sRM[k + 8];
						uint32_t i0 = (uint32_t)sRM[16 + k], i1 = (uint32_t)sRM[16 + k + 8];
						if(sRM[k + 4] > m0){ m0 = sRM[k + 4]; i0 = (uint32_t)sRM[16 + k + 4]; }
						if(sRM[k + 12] > m1){ m1 = sRM[k + 12]; i1 = (uint32_t)sRM[16 + k + 12]; }
						if(m1 > m0){ m0 = m1; i0 = i1; }
						M4[k] = m0; I4[k] = i0;
					}
					int max_score = M4[0]; uint32_t bi = I4[0];
					#pragma unroll
					for(int k=1;k<4;k++) if(M4[k] > max_score){ max_score = M4[k]; bi = I4[k]; }
					if(max_score > maxscr){
						const uint32_t bl = bi & 0xff;
						uint32_t pos = 0; int umax = kScoreMin, uscr = 0;
						for(uint32_t x=0;x<W;x++){ uscr += sLin[bl * 8 + x]; if(uscr > umax){ pos = x; umax = uscr; } }
						maxscr = max_score; maxidx = (int)u; maxoff = (int)(bl * W + pos) + rpos_u;
					}
					__syncwarp(gmask);
				}

				vst[v] = vst[v] + 1;
				continue;
			}
			if(sp){
				// pop: the top of the stack is in a register, deeper entries in HBM scratch
				u = tos;
				sp--;
				if(sp) tos = stack[sp - 1];
				const int2 un = node[u];
				rpos_u = un.x; base_u = ((uint32_t)un.y >> 16) & 0xffu;
				mpos_u = (mode != 1) ? mpos[u] : 0;
				ei = eoff[u]; e1 = eoff[u + 1];
				if(cur_node == u) Pd = Cu;
				else row8_load<PW>(Pd, rows + (size_t)u * mmblk, t);
				continue;
			}
			// the job is finished
			if(t == 0){
				a.best[(size_t)job * 3 + 0] = maxscr; a.best[(size_t)job * 3 + 1] = maxidx; a.best[(size_t)job * 3 + 2] = maxoff;
				a.status[job] = stflag;
				a.ops[(size_t)job * 2 + 0] = nupd; a.ops[(size_t)job * 2 + 1] = nmrg;
			}
			have = false;
		}
		if(__all_sync(amask, done)) break;
		// =========================== phase B: one row update (+ merge) per group ==========================================
		if(act){
			// ================= dpalign_row_update_bspoa (bspoa.h:2232-2261) =================================
			const int2 vn = node[v];
			const int rpos_v = vn.x;
			const uint32_t nct_v = (uint32_t)vn.y & 0xffffu, base_v = ((uint32_t)vn.y >> 16) & 0xffu, bonus_v = ((uint32_t)vn.y >> 24) & 1u;
			const uint32_t vst_v = vst[v];
			const uint32_t m = (uint32_t)rpos_v - (uint32_t)rpos_u;
			const uint32_t q1 = (uint32_t)rpos_u, q2 = (uint32_t)rpos_v;
			// ---- substitution scores of this edge for the thread's 16 cells (bspoa.h:2588, 2199-2215) ----
			uint32_t z0, z1, z2, z3, z4, z5, z6, z7;
			{
				const int Mk = (int8_t)(Mm + (bonus_v ? refbonus : 0)), Xk = (int8_t)Xx;
				const int hpc = (base_v != base_u) ? 1 : 0;
				const uint32_t bsh = 8u * (base_v & 3u);
				const uint32_t Tlo = ((0x01010101u * (uint32_t)(Xk & 0xff)) & ~(0xffu << bsh)) | ((uint32_t)(Mk & 0xff) << bsh);
				const uint32_t Thi = ((0x01010101u * (uint32_t)((Xk + hpc) & 0xff)) & ~(0xffu << bsh)) | ((uint32_t)((Mk + hpc) & 0xff) << bsh);
				const uint32_t x0 = q2 + 16u * (uint32_t)t;
				const uint32_t *wp = qsel + (x0 >> 2);
				const uint32_t sh = (x0 & 3u) * 8u;
				const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
				const uint32_t s0 = __funnelshift_r(w0, w1, sh), s1 = __funnelshift_r(w1, w2, sh), s2 = __funnelshift_r(w2, w3, sh), s3 = __funnelshift_r(w3, w4, sh);
				constexpr uint32_t BIAS = 0x00800080u;
				z0 = __viaddmax_s16x2(prmt(Tlo, Thi, prmt(s0, s2, 0x0040u)), BIAS, Z0);
				z1 = __viaddmax_s16x2(prmt(Tlo, Thi, prmt(s0, s2, 0x0051u)), BIAS, Z0);
				z2 = __viaddmax_s16x2(prmt(Tlo, Thi, prmt(s0, s2, 0x0062u)), BIAS, Z0);
				z3 = __viaddmax_s16x2(prmt(Tlo, Thi, prmt(s0, s2, 0x0073u)), BIAS, Z0);
				z4 = __viaddmax_s16x2(prmt(Tlo, Thi, prmt(s1, s3, 0x0040u)), BIAS, Z0);
				z5 = __viaddmax_s16x2(prmt(Tlo, Thi, prmt(s1, s3, 0x0051u)), BIAS, Z0);
				z6 = __viaddmax_s16x2(prmt(Tlo, Thi, prmt(s1, s3, 0x0062u)), BIAS, Z0);
				z7 = __viaddmax_s16x2(prmt(Tlo, Thi, prmt(s1, s3, 0x0073u)), BIAS, Z0);
			}
			// ---- the predecessor row shifted to v's band (bsalign.h:2244-2392) ----
			uint32_t pu0 = Pd.u0, pu1 = Pd.u1, pu2 = Pd.u2, pu3 = Pd.u3;
			uint32_t pe0 = Pd.e0, pe1 = Pd.e1, pe2 = Pd.e2, pe3 = Pd.e3;
			uint32_t pq0 = Pd.q0, pq1 = Pd.q1, pq2 = Pd.q2, pq3 = Pd.q3;
			int ubsA = Pd.ubA, ubsB = Pd.ubB, ubs16 = Pd.ub16;
			if(m != 0){
				if(m >= bw){
					pu0 = pu1 = pu2 = pu3 = 0x80808080u; pe0 = pe1 = pe2 = pe3 = 0u; pq0 = pq1 = pq2 = pq3 = 0u;
					ubsA = ubsB = ubs16 = kScoreMin;
				} else if(m <= 4){
					// anchors advance by the lane's first m cells; only the band end crosses into the overhang
					uint32_t l0, l1, l2, l3;
					deinterleave(pu0, pu1, pu2, pu3, l0, l1, l2, l3);
					const uint32_t wm = 0x01010101u >> (8u * (4u - m));
					ubsA += (int)__dp4a(l0, wm, 0u) - 128 * (int)m;
					ubsB += (int)__dp4a(l2, wm, 0u) - 128 * (int)m;
					{
						const uint32_t n1 = (m - 1 < ovd - 1) ? m - 1 : ovd - 1;
						ubs16 += ovc + (int)n1 * ge1 + (int)(m - 1 - n1) * ge2;
					}
					const uint32_t r8 = 8u * m;
					slide4(pu0, pu1, pu2, pu3, r8, gmask, t, over_u);
					slide4(pe0, pe1, pe2, pe3, r8, gmask, t, 0u);
					if(PW == 2) slide4(pq0, pq1, pq2, pq3, r8, gmask, t, 0u);
				} else {
					// general shift (rare): through shared memory, cell by cell
					uint32_t l0, l1, l2, l3;
					__syncwarp(gmask);
					deinterleave(pu0, pu1, pu2, pu3, l0, l1, l2, l3); *(uint4*)(sLin + 16 * t) = make_uint4(l0, l1, l2, l3);
					deinterleave(pe0, pe1, pe2, pe3, l0, l1, l2, l3); *(uint4*)(sLin + 128 + 16 * t) = make_uint4(l0, l1, l2, l3);
					deinterleave(pq0, pq1, pq2, pq3, l0, l1, l2, l3); *(uint4*)(sLin + 256 + 16 * t) = make_uint4(l0, l1, l2, l3);
					sUB[A] = ubsA; sUB[B] = ubsB; if(t == 7) sUB[16] = ubs16;
					__syncwarp(gmask);
					const uint32_t cyc = m / W, mr = m - cyc * W, i0 = bw - m;
					auto anchor = [&](uint32_t k) -> int {
						int s;
						if(k + cyc < (uint32_t)kLanes){
							s = sUB[k + cyc];
							for(uint32_t x=0;x<mr;x++) s += (int)(uint8_t)sLin[(k + cyc) * 8 + x] - 128;
						} else s = sUB[kLanes];
						const uint32_t Pp = k * W;
						if(k >= 1 && Pp > i0){
							const uint32_t kk = Pp - i0;
							const uint32_t n1 = (kk - 1 < ovd - 1) ? kk - 1 : ovd - 1;
							s += ovc + (int)n1 * ge1 + (int)(kk - 1 - n1) * ge2;
						}
						return s;
					};
					ubsA = anchor((uint32_t)A); ubsB = anchor((uint32_t)B); ubs16 = anchor(16u);
					uint32_t nu[4] = {0, 0, 0, 0}, ne[4] = {0, 0, 0, 0}, nq[4] = {0, 0, 0, 0};
					#pragma unroll
					for(int b=0;b<16;b++){
						const uint32_t p = 16u * (uint32_t)t + m + (uint32_t)b;
						uint32_t uv, ev = 0, qv = 0;
						if(p < bw){ uv = (uint8_t)sLin[p]; ev = (uint8_t)sLin[128 + p]; qv = (uint8_t)sLin[256 + p]; }
						else uv = (uint32_t)((POA_OVER(p - bw) + 128) & 0xff);
						nu[b >> 2] |= uv << (8 * (b & 3)); ne[b >> 2] |= ev << (8 * (b & 3)); nq[b >> 2] |= qv << (8 * (b & 3));
					}
					interleave(nu[0], nu[1], nu[2], nu[3], pu0, pu1, pu2, pu3);
					interleave(ne[0], ne[1], ne[2], ne[3], pe0, pe1, pe2, pe3);
					interleave(nq[0], nq[1], nq[2], nq[3], pq0, pq1, pq2, pq3);
					__syncwarp(gmask);
				}
			}
			// ---- cell 0 of the band (bsalign.h:2899-2907), thread 0 ----
			uint32_t zmask = 0xffffffffu, zor = 0u;
			if(t == 0){
				int rh;
				if(q1 == q2){
					if(q1) rh = kScoreMin;
					else if(mode == 1 || mpos_v == 0) rh = 0;
					else if(PW < 2) rh = (int)((uint32_t)O + (uint32_t)E * (uint32_t)mpos_v);
					else { uint32_t c1 = (uint32_t)O + (uint32_t)E * (uint32_t)mpos_v, c2 = (uint32_t)Q + (uint32_t)P * (uint32_t)mpos_v; rh = (int)(c1 > c2 ? c1 : c2); }
				} else if(q1 + bw >= q2) rh = ubsA;
				else rh = kScoreMin;
				const int zc = lo16(z0) - 128;
				const int up = (int)(pu0 & 0xffu) - 128, ep = (int)(int8_t)(pe0 & 0xffu), qp = (int)(int8_t)(pq0 & 0xffu);
				int h0 = (rh - ubsA) + zc;
				const int t0 = up + (PW == 2 ? max(ep, qp) : ep);
				if(h0 >= t0){ if(h0 > kEpi8Max) h0 = kEpi8Max; } else h0 = kEpi8Min;
				zmask = 0xffff0000u; zor = (uint32_t)((h0 + 128) & 0xffff);
			}
			z0 = (z0 & zmask) | zor;
			const uint4 cu4 = make_uint4(pu0, pu1, pu2, pu3), ce4 = make_uint4(pe0, pe1, pe2, pe3), cq4 = make_uint4(pq0, pq1, pq2, pq3);
			// ---- pass 1: F (G) leaving the two lanes' running blocks with nothing entering ----
			RowState st; st.f = pk1(kEpi8Min + 128); st.g = pk1(kEpi8Min + 128); st.h = 0; st.u = 0; st.nv = 0;
			{
				uint32_t d0, d1, d2;
				#define P1(K, Z) dp_step<PW, true, false>(st, entz<K>(cu4), ent<K>(ce4), ent<K>(cq4), Z, dpk, d0, d1, d2);
				P1(0, z0) P1(1, z1) P1(2, z2) P1(3, z3) P1(4, z4) P1(5, z5) P1(6, z6) P1(7, z7)
				#undef P1
			}
			// ---- F penetration (bsalign.h:2639-2652) ----
			{
				const int fendA = lo16(st.f) - 128, fendB = hi16(st.f) - 128;
				const int gendA = lo16(st.g) - 128, gendB = hi16(st.g) - 128;
				const int tW = (int)W * ge1, tW2 = (int)W * ge2;
				int nxt = __shfl_down_sync(gmask, ubsA, 1, kPoaFastGroup);
				if(t == 7) nxt = ubs16;
				const int dA = ubsB - ubsA, dB = nxt - ubsB;
				// guess: every lane's entry F is the exit F of the lane before it; valid iff no lane's chain value exceeds it
				int finA = __shfl_up_sync(gmask, fendB, 1, kPoaFastGroup), finB = fendA;
				int ginA = __shfl_up_sync(gmask, gendB, 1, kPoaFastGroup), ginB = gendA;
				if(t == 0){ finA = kEpi8Min; ginA = kEpi8Min; }
				const int sB = tW + finA - dA, sB2 = tW2 + ginA - dA;
				int sA = __shfl_up_sync(gmask, tW + finB - dB, 1, kPoaFastGroup), sA2 = __shfl_up_sync(gmask, tW2 + ginB - dB, 1, kPoaFastGroup);
				bool ok = (finB >= sB) && (t == 0 || finA >= sA);
				if(PW == 2) ok = ok && (ginB >= sB2) && (t == 0 || ginA >= sA2);
				if(!__all_sync(gmask, ok)){
					// exact 16-step scan, every thread redundantly
					sF[A] = (int8_t)fendA; sF[B] = (int8_t)fendB; sF[16 + A] = (int8_t)gendA; sF[16 + B] = (int8_t)gendB;
					sD[A] = dA; sD[B] = dB;
					__syncwarp(gmask);
					int s = tW + kEpi8Min - sD[0], s2 = tW2 + kEpi8Min - sD[0];
					#pragma unroll
					for(int k=1;k<kLanes;k++){
						int fk = sF[k - 1];
						if(fk < s) fk = (int)(int8_t)s;
						int gk = 0;
						if(PW == 2){ gk = sF[16 + k - 1]; if(gk < s2) gk = (int)(int8_t)s2; }
						if(k == A){ finA = fk; ginA = gk; }
						if(k == B){ finB = fk; ginB = gk; }
						s = tW + fk - sD[k];
						if(PW == 2) s2 = tW2 + gk - sD[k];
					}
					__syncwarp(gmask);
				}
				st.f = pk(finA + 128, finB + 128); st.g = pk(ginA + 128, ginB + 128);
			}
			// ---- pass 2: the row (bsalign.h:2934-2957, 3141-3176) ----
			st.nv = 0; st.h = 0; st.u = 0;
			uint32_t un0, un1, un2, un3, un4, un5, un6, un7, en0 = 0, en1 = 0, en2 = 0, en3 = 0, en4 = 0, en5 = 0, en6 = 0, en7 = 0,
				qn0 = 0, qn1 = 0, qn2 = 0, qn3 = 0, qn4 = 0, qn5 = 0, qn6 = 0, qn7 = 0;
			#define P2(K, Z) dp_step<PW, true, true>(st, entz<K>(cu4), ent<K>(ce4), ent<K>(cq4), Z, dpk, un##K, en##K, qn##K);
			P2(0, z0) P2(1, z1) P2(2, z2) P2(3, z3) P2(4, z4) P2(5, z5) P2(6, z6) P2(7, z7)
			#undef P2
			// ---- tail (bsalign.h:2618-2636) ----
			{
				constexpr int HB = 257;   // FAST keeps h biased by 128 + 129 behind the gap-open add
				uint32_t h = pk(lo16(st.h) - HB, hi16(st.h) - HB);
				const uint32_t ul = pk(lo16(st.u) - 128, hi16(st.u) - 128);
				if(PW == 1) h = sadd(h, NGOE); else h = sadd(h, NGQP);
				const uint32_t vt = ssubc(h, ~ul);
				const int vtA = lo16(vt), vtB = hi16(vt);
				int vprev = __shfl_up_sync(gmask, vtB, 1, kPoaFastGroup);
				if(t == 0) vprev = 0;
				int uA = clamp8(lo16(un0) - 128 - vprev);
				const int uB = clamp8(hi16(un0) - 128 - vtA);
				Cu.ubB = ubsB + vtA;
				Cu.ub16 = ubs16 + vtB;
				if(t == 0){ Cu.ubA = ubsA + uA; uA = 0; }
				else Cu.ubA = ubsA + vprev;
				un0 = pk(uA + 128, uB + 128);
			}
			Cu.u0 = pack2(un0, un1); Cu.u1 = pack2(un2, un3); Cu.u2 = pack2(un4, un5); Cu.u3 = pack2(un6, un7);
			Cu.e0 = pack2(en0, en1); Cu.e1 = pack2(en2, en3); Cu.e2 = pack2(en4, en5); Cu.e3 = pack2(en6, en7);
			if(PW == 2){ Cu.q0 = pack2(qn0, qn1); Cu.q1 = pack2(qn2, qn3); Cu.q2 = pack2(qn4, qn5); Cu.q3 = pack2(qn6, qn7); }
			else { Cu.q0 = Cu.q1 = Cu.q2 = Cu.q3 = 0u; }
			nupd++;
			uint8_t *vblk = rows + (size_t)v * mmblk;
			// ---- later visits: merge into the block of v (bspoa.h:2263-2272, bsalign.h:2474-2616) ----
			if(vst_v){
				Row8 Bo;
				row8_load<PW>(Bo, vblk, t);
				// lanes never saturate int16 within 8 steps (|start| <= 16384, |cell| <= 128), so plain s16x2 arithmetic is exact
				auto half = [](int d) -> int { d = max(-0x7FFF, min(0x7FFF, d)); return d >> 1; };
				const int dAa = max(-0x7FFF, min(0x7FFF, Cu.ubA - Bo.ubA)), dBb = max(-0x7FFF, min(0x7FFF, Cu.ubB - Bo.ubB));
				const int xaA = half(Cu.ubA - Bo.ubA), xaB = half(Cu.ubB - Bo.ubB);
				uint32_t ta = pk(xaA, xaB), tb = pk(xaA - dAa, xaB - dBb);
				uint32_t mp = __vmaxs2(ta, tb);
				const uint4 au4 = make_uint4(Cu.u0 ^ 0x80808080u, Cu.u1 ^ 0x80808080u, Cu.u2 ^ 0x80808080u, Cu.u3 ^ 0x80808080u);
				const uint4 bu4 = make_uint4(Bo.u0 ^ 0x80808080u, Bo.u1 ^ 0x80808080u, Bo.u2 ^ 0x80808080u, Bo.u3 ^ 0x80808080u);
				const uint4 ae4 = make_uint4(Cu.e0, Cu.e1, Cu.e2, Cu.e3), be4 = make_uint4(Bo.e0, Bo.e1, Bo.e2, Bo.e3);
				const uint4 aq4 = make_uint4(Cu.q0, Cu.q1, Cu.q2, Cu.q3), bq4 = make_uint4(Bo.q0, Bo.q1, Bo.q2, Bo.q3);
				constexpr uint32_t NOLIM = 0x80008000u;   // per-half add whose lower bound never binds
				uint32_t mu0, mu1, mu2, mu3, mu4, mu5, mu6, mu7, me0, me1, me2, me3, me4, me5, me6, me7,
					mq0 = 0, mq1 = 0, mq2 = 0, mq3 = 0, mq4 = 0, mq5 = 0, mq6 = 0, mq7 = 0;
				#define MG(K) { \
					ta = __viaddmax_s16x2(ta, ent<K>(au4), NOLIM); tb = __viaddmax_s16x2(tb, ent<K>(bu4), NOLIM); \
					const uint32_t mc = __vmaxs2(ta, tb), nmc = ~mc; \
					mu##K = ssubc(mc, ~mp); mp = mc; \
					me##K = ssubc(__vmaxs2(__viaddmax_s16x2(ta, ent<K>(ae4), NOLIM), __viaddmax_s16x2(tb, ent<K>(be4), NOLIM)), nmc); \
					if(PW == 2) mq##K = ssubc(__vmaxs2(__viaddmax_s16x2(ta, ent<K>(aq4), NOLIM), __viaddmax_s16x2(tb, ent<K>(bq4), NOLIM)), nmc); }
				MG(0) MG(1) MG(2) MG(3) MG(4) MG(5) MG(6) MG(7)
				#undef MG
				Cu.u0 = pack2(mu0, mu1) ^ 0x80808080u; Cu.u1 = pack2(mu2, mu3) ^ 0x80808080u; Cu.u2 = pack2(mu4, mu5) ^ 0x80808080u; Cu.u3 = pack2(mu6, mu7) ^ 0x80808080u;
				Cu.e0 = pack2(me0, me1); Cu.e1 = pack2(me2, me3); Cu.e2 = pack2(me4, me5); Cu.e3 = pack2(me6, me7);
				if(PW == 2){ Cu.q0 = pack2(mq0, mq1); Cu.q1 = pack2(mq2, mq3); Cu.q2 = pack2(mq4, mq5); Cu.q3 = pack2(mq6, mq7); }
				Cu.ubA = max(Cu.ubA, Bo.ubA); Cu.ubB = max(Cu.ubB, Bo.ubB); Cu.ub16 = max(Cu.ub16, Bo.ub16);
				nmrg++;
			}

			row8_store<PW>(Cu, vblk, t);
			cur_node = v;
			vst[v] = vst_v + 1;
			if(vst_v + 1 == nct_v){
				if(mode != 0 && q2 + bw >= slen){
					int smx = getscore(Cu, (int64_t)slen - 1 - (int64_t)q2) + T;
					if(smx > maxscr){ maxscr = smx; maxidx = (int)v; maxoff = (int)slen - 1; }
				}
				// push: the previous top goes to HBM scratch
				if(sp) stack[sp - 1] = tos;
				tos = v; sp++;
			}
		}
	}
	#undef POA_OVER
}

} // namespace bsb200
