// edit_kernels.cuh -- sm_100a kernel for the 2-bit-plane edit-distance DP (replaces bsalign.h:612-1206).
//
// The reference keeps u(x,y) = H(x,y) - H(x-1,y) in {-1,0,+1} as two bit-planes (minus, plus) striped over
// the 64 bit-lanes of uint64 words and resolves the left-to-right dependency of a row with a re-pass loop
// that runs to the exact fix-point (bsalign.h:784-809).  That fix-point is the sequential recurrence
//     h = (q[x] != t[y]) & (u(x,y-1) != -1) & (v(x-1,y) != -1);  u' = h - v;  v' = h - u
// so here a row is evaluated in LINEAR bit order with the Myers/Hyyro carry trick (one 64-bit add per word
// resolves the whole horizontal chain), which yields the identical planes without any re-pass.
//
// Mapping: one THREAD per pair (a 300x300 pair at band 64 is one machine word per row), 32 pairs of similar
// length per warp, rows of the 32 pairs interleaved in the HBM trace so that every warp-wide row store is
// one contiguous 256-byte line per plane.  Forward sweep and backtrace are fused in one kernel, which saves a launch and the
// per-pair state, not the traffic: at configs[3]'s size the interleaved trace of all resident warps (> 1 GB per 200k pairs) does not
// stay in the 126 MB L2, so the walk reads it back from HBM once (ncu: 1.05 GB written + 1.08 GB read per 200k pairs, DRAM / algorithmic
// bytes 1.76; profiles/ncu_full_edit_c4_200kpairs_r2.txt).
#pragma once
#include "common.cuh"
#include <stdint.h>

namespace bsb200 {

constexpr int kEditThreads = 128;

// bsalign.h:1055-1067
__host__ __device__ inline uint32_t edit_bandwidth(uint32_t qlen, uint32_t tlen, int mode, uint32_t bandwidth){
	uint32_t q64 = (qlen + 63) / 64 * 64, bw;
	int type = mode & 3;
	if(type == 1 || type == 2) return q64;
	bw = (bandwidth + 63) / 64 * 64;
	if(bw == 0 || bw > qlen) bw = q64;
	if(tlen && bw < qlen && bw < ((qlen + tlen - 1) / tlen) + 1) bw = ((qlen + tlen - 1) / tlen + 1 + 63) / 64 * 64;
	return bw;
}

// bytes of interleaved trace one pair contributes when every pair of its 32-block has the block's row count
__host__ __device__ inline uint64_t edit_trace_bytes(uint32_t bwmax, uint32_t tlen){
	return ((uint64_t)(bwmax / 64) * 16 + 4) * ((uint64_t)tlen + 1);
}

struct EditArgs {
	const uint8_t *seqs;
	const uint64_t *qoff, *toff;
	const uint32_t *qlen, *tlen;
	const uint32_t *order;       // pairs of this wave, longest target first; 32 consecutive entries form a block
	uint32_t npairs;
	uint8_t *trace;
	const uint64_t *block_off;   // per 32-block byte offset into trace
	const uint32_t *block_rows;  // per 32-block row count (max tlen in block + 1)
	int32_t *results, *status;
	uint32_t *cigars; const uint64_t *cig_off;
	uint32_t *dense; uint64_t *dense_off; unsigned long long *dense_total;
	uint32_t *ncigar;
	int mode;
	uint32_t bandwidth;
	uint32_t WB;                 // words per row in the trace layout (batch maximum)
	uint32_t nQW;                // query bit-plane words per thread in shared memory
};

__device__ __forceinline__ uint64_t lowmask64(uint32_t n){ return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

template<int WR>
__global__ void __launch_bounds__(kEditThreads) edit_kernel(const EditArgs a){
	constexpr bool REG = WR <= 16;
	extern __shared__ __align__(16) uint64_t qbits[];   // [word][plane][thread]
	const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t nthr = blockDim.x;
	if(gid >= a.npairs) return;
	const uint32_t pair = a.order[gid];
	const uint32_t qlen = a.qlen[pair], tlen = a.tlen[pair];
	const uint8_t *qs = a.seqs + a.qoff[pair], *ts = a.seqs + a.toff[pair];
	int32_t *rs = a.results + (size_t)pair * 10;
	const int type = a.mode & 3;
	const uint32_t bw = edit_bandwidth(qlen, tlen, type, a.bandwidth);
	const uint32_t q64 = (qlen + 63) / 64 * 64;
	const uint32_t W = bw / 64;
	int err = 0;
	if(bw > q64 || W > (uint32_t)WR){ a.status[pair] = 1; if(a.ncigar) a.ncigar[pair] = 0; if(a.dense_off) a.dense_off[pair] = 0; return; }
	// ---- trace block of this warp's 32 pairs --------------------------------------------------------
	const uint32_t blk = gid >> 5;
	const uint32_t R = a.block_rows[blk];
	uint64_t *tp = (uint64_t*)(a.trace + a.block_off[blk]);                 // planes: ((row*WB + w)*2 + plane)*32 + lane
	(void)R;
	#define TP(row, w, plane) tp[(((size_t)(row) * a.WB + (w)) * 2 + (plane)) * 32 + lane]
	// band offset of target row y (bsalign.h:1112-1114): a function of y alone, so the backtrace computes it instead of reading it back
	auto rbeg_of = [&](int y) -> uint32_t {
		if(type != 0) return 0u;
		uint32_t r_ = (uint32_t)(((uint64_t)(uint32_t)y * qlen) / tlen);
		r_ = (r_ < bw / 2) ? 0u : r_ - bw / 2;
		return (r_ + bw > q64) ? q64 - bw : r_;
	};
	#define RBEG(y) rbeg_of(y)
	// ---- query bit-planes in shared memory (bit x of plane 0/1 = low/high bit of q[x]) --------------
	#define QB(word, plane) qbits[((size_t)(word) * 2 + (plane)) * nthr + threadIdx.x]
	const uint32_t nqw = a.nQW;
	for(uint32_t w=0;w<nqw;w++){
		uint64_t lo = 0, hi = 0;
		uint32_t x0 = w * 64;
		if(x0 < qlen){
			uint32_t n = qlen - x0 < 64 ? qlen - x0 : 64;
			for(uint32_t k=0;k<n;k++){ uint32_t c = qs[x0 + k]; lo |= (uint64_t)(c & 1) << k; hi |= (uint64_t)((c >> 1) & 1) << k; }
		}
		QB(w, 0) = lo; QB(w, 1) = hi;
	}
	// ---- row state ----------------------------------------------------------------------------------
	uint64_t Pv[WR], Mv[WR];
	#pragma unroll (REG ? WR : 1)
	for(int w=0;w<WR;w++){ if(!REG && w >= (int)W) break; Pv[w] = ~0ull; Mv[w] = 0ull; }
	for(uint32_t w=0;w<W;w++){ TP(0, w, 0) = 0ull; TP(0, w, 1) = ~0ull; }
	int sbeg = 0, smin = 0x7FFFFFFF, rx = (int)qlen - 1, ry = (int)tlen - 1;
	uint32_t rbeg = 0, prev_beg = 0;
	const uint32_t qd = qlen / tlen, qr = qlen % tlen;   // floor(i*qlen/tlen) kept incrementally (bsalign.h:1112)
	uint32_t dq = 0, dr = 0;
	const uint64_t hin = (type == 1) ? 0ull : 1ull;
	for(uint32_t i=0;i<tlen;i++){
		const uint32_t tbase = ts[i];
		if(type == 0){
			rbeg = dq;
			rbeg = (rbeg < bw / 2) ? 0 : rbeg - bw / 2;
			if(rbeg + bw > q64) rbeg = q64 - bw;
			dq += qd; dr += qr; if(dr >= tlen){ dr -= tlen; dq++; }
		} else rbeg = 0;
		const uint32_t movx = rbeg - prev_beg;
		// ---- shift the previous row to this band (bsalign.h:658-721) -------------------------------
		if(type == 1) sbeg = 0;
		else {
			if(movx){
				if(movx < bw && movx >= W && movx % W) err |= 8; // the reference corrupts its scratch here (bsalign.h:704-713)
				uint32_t m = movx < bw ? movx : bw, w = 0;
				// sbeg follows H(rbeg-1, y): add the u of the cells the band leaves behind
				#pragma unroll (REG ? WR : 1)
				for(int k=0;k<WR;k++){
					if(!REG && k >= (int)W) break;
					if((uint32_t)k < W){
						uint32_t lo = (uint32_t)k * 64, n = m > lo ? (m - lo < 64 ? m - lo : 64) : 0;
						uint64_t mk = lowmask64(n);
						sbeg += __popcll(Pv[k] & mk) - __popcll(Mv[k] & mk);
					}
				}
				(void)w;
				if(movx >= bw){
					#pragma unroll (REG ? WR : 1)
					for(int k=0;k<WR;k++){ if(!REG && k >= (int)W) break; Pv[k] = ~0ull; Mv[k] = 0ull; }
				} else {
					uint32_t ws = movx / 64, bs = movx % 64;
					if(REG){
						for(;ws;ws--){ // whole-word moves with static indices keep the planes in registers
							#pragma unroll
							for(int k=0;k<WR;k++){
								bool last = ((uint32_t)k + 1 >= W);
								Pv[k] = (k + 1 < WR && !last) ? Pv[k + 1 < WR ? k + 1 : k] : ~0ull;
								Mv[k] = (k + 1 < WR && !last) ? Mv[k + 1 < WR ? k + 1 : k] : 0ull;
							}
						}
						if(bs){
							#pragma unroll
							for(int k=0;k<WR;k++){
								bool last = ((uint32_t)k + 1 >= W);
								uint64_t pn = (k + 1 < WR && !last) ? Pv[k + 1 < WR ? k + 1 : k] : ~0ull;
								uint64_t mn = (k + 1 < WR && !last) ? Mv[k + 1 < WR ? k + 1 : k] : 0ull;
								Pv[k] = (Pv[k] >> bs) | (pn << (64 - bs));
								Mv[k] = (Mv[k] >> bs) | (mn << (64 - bs));
							}
						}
					} else {
						for(uint32_t k=0;k<W;k++){
							uint64_t p0 = (k + ws < W) ? Pv[k + ws] : ~0ull, p1 = (k + ws + 1 < W) ? Pv[k + ws + 1] : ~0ull;
							uint64_t m0 = (k + ws < W) ? Mv[k + ws] : 0ull, m1 = (k + ws + 1 < W) ? Mv[k + ws + 1] : 0ull;
							Pv[k] = bs ? (p0 >> bs) | (p1 << (64 - bs)) : p0;
							Mv[k] = bs ? (m0 >> bs) | (m1 << (64 - bs)) : m0;
						}
					}
				}
			}
			sbeg++;
		}
		// ---- the row: Myers/Hyyro step over W words, carries chained --------------------------------
		const uint64_t TL = (tbase & 1) ? ~0ull : 0ull, TH = (tbase & 2) ? ~0ull : 0ull;
		const uint32_t qw0 = rbeg >> 6, qsh = rbeg & 63;
		uint64_t carry = 0, phin = hin, mhin = 0;
		int rowsum = 0;
		#pragma unroll (REG ? WR : 1)
		for(int w=0;w<WR;w++){
			if((uint32_t)w >= W) break;
			// match mask of query positions rbeg + 64w .. +63 against tbase
			uint64_t l0 = QB(qw0 + w, 0), h0 = QB(qw0 + w, 1), l1 = QB(qw0 + w + 1, 0), h1 = QB(qw0 + w + 1, 1);
			uint64_t ql = qsh ? (l0 >> qsh) | (l1 << (64 - qsh)) : l0;
			uint64_t qh = qsh ? (h0 >> qsh) | (h1 << (64 - qsh)) : h0;
			uint32_t x0 = rbeg + 64 * (uint32_t)w;
			uint64_t valid = x0 >= qlen ? 0ull : lowmask64(qlen - x0);
			uint64_t Eq = ~(ql ^ TL) & ~(qh ^ TH) & valid;
			uint64_t pv = Pv[w], mv = Mv[w];
			uint64_t Xv = Eq | mv;
			uint64_t a0 = Eq & pv;
			uint64_t sum = a0 + pv;
			uint64_t c1 = sum < a0;
			uint64_t sum2 = sum + carry;
			c1 |= (sum2 < sum);
			carry = c1;
			uint64_t Xh = (sum2 ^ pv) | Eq;
			uint64_t Ph = mv | ~(Xh | pv);
			uint64_t Mh = pv & Xh;
			uint64_t pho = Ph >> 63, mho = Mh >> 63;
			Ph = (Ph << 1) | phin; Mh = (Mh << 1) | mhin;
			phin = pho; mhin = mho;
			pv = Mh | ~(Xv | Ph);
			mv = Ph & Xv;
			Pv[w] = pv; Mv[w] = mv;
			TP(i + 1, w, 0) = mv; TP(i + 1, w, 1) = pv;
			rowsum += __popcll(pv & valid) - __popcll(mv & valid);
		}
		if(type != 0){ // bsalign.h:1124-1139: H(qlen-1, i)
			int srow = sbeg + rowsum;
			if(srow < smin){ smin = srow; rx = (int)qlen - 1; ry = (int)i; }
		} else if(i + 1 == tlen){
			smin = sbeg + rowsum; // global score (bsalign.h:1194-1202)
		}
		prev_beg = rbeg;
	}
	// ---- EXTEND: arg-min over the last row with the reference's lane/chunk order (bsalign.h:813-963) ---
	if(type == 2){
		int sb = sbeg, best = sbeg; uint32_t pmin = 0;
		for(uint32_t blk4=0;blk4<4;blk4++){
			int cand[16], tot[16]; uint32_t pp[16];
			#pragma unroll 1
			for(uint32_t l=0;l<16;l++){
				uint32_t j = blk4 * 16 + l; // bit-lane j covers band positions [j*W, (j+1)*W)
				int hh = 0, mm = 0; uint32_t ppos = 0;
				for(uint32_t ib=0;ib<W;ib+=124){
					uint32_t ie = ib + 124 < W ? ib + 124 : W;
					int h = 0, m = 0; uint32_t pz = 0;
					for(uint32_t k=ib;k<ie;k++){
						uint32_t p = j * W + k;
						uint64_t pw_ = TP(tlen, p >> 6, 1), mw_ = TP(tlen, p >> 6, 0);
						h += (int)((pw_ >> (p & 63)) & 1) - (int)((mw_ >> (p & 63)) & 1);
						if(m > h){ m = h; pz = k - ib; }
					}
					int d = hh + m;
					if(mm > d){ mm = d; ppos = pz + ib; }
					hh += h;
				}
				tot[l] = hh; cand[l] = mm; pp[l] = ppos;
			}
			int sc = 0; uint32_t st = 0;
			for(uint32_t l=0;l<16;l++){ int c = sb + cand[l]; sb += tot[l]; if(l == 0 || sc > c){ sc = c; st = l; } }
			if(sc >= best) continue;
			best = sc; pmin = (blk4 * 16 + st) * W + pp[st];
		}
		if(best < smin){ smin = best; rx = (int)pmin; ry = (int)tlen - 1; }
	}
	// ---- backtrace (bsalign.h:965-1044) -------------------------------------------------------------------
	CigarSink cg;
	cg.buf = a.cigars ? a.cigars + a.cig_off[pair] : nullptr;
	cg.cap = a.cigars ? (uint32_t)(a.cig_off[pair + 1] - a.cig_off[pair]) : 0;
	cg.n = 0; cg.run = 0; cg.err = 0;
	int x = rx, y = ry, mat = 0, mis = 0, ins = 0, del = 0;
	const int qe = x + 1, te = y + 1;
	int64_t guard = 0;
	while(x >= 0 && y >= 0){
		if(++guard > 4 * ((int64_t)qlen + tlen) + 64){ err |= 2; break; }
		uint32_t qc = (uint32_t)((QB(x >> 6, 0) >> (x & 63)) & 1) | (uint32_t)(((QB(x >> 6, 1) >> (x & 63)) & 1) << 1);
		uint32_t op;
		if(qc == ts[y]){ mat++; op = 0; x--; y--; }
		else {
			// both rows the decision may need are requested together (slot y + 1 = row y, slot y = row y - 1; slot 0 is the init row)
			const int64_t p1 = (int64_t)x - (int64_t)RBEG(y), p0 = (int64_t)x - (int64_t)(y > 0 ? RBEG(y - 1) : 0u);
			const bool ok1 = p1 >= 0 && p1 < (int64_t)bw, ok0 = p0 >= 0 && p0 < (int64_t)bw;
			const uint64_t w1p = ok1 ? TP(y + 1, p1 >> 6, 1) : 0ull, w1m = ok1 ? TP(y + 1, p1 >> 6, 0) : 0ull;
			const uint64_t w0p = ok0 ? TP(y, p0 >> 6, 1) : 0ull, w0m = ok0 ? TP(y, p0 >> 6, 0) : 0ull;
			int u_here = 0;
			if(!ok1) err |= 1;
			else u_here = (int)((w1p >> (p1 & 63)) & 1) - (int)((w1m >> (p1 & 63)) & 1);
			if(u_here == 1){ ins++; op = 1; x--; }
			else {
				int u_up = 0;
				if(!ok0) err |= 1;
				else u_up = (int)((w0p >> (p0 & 63)) & 1) - (int)((w0m >> (p0 & 63)) & 1);
				if(u_up == -1){ del++; op = 2; y--; }
				else { mis++; op = 0; x--; y--; }
			}
		}
		if(op == (cg.run & 0xf)) cg.run += 0x10;
		else { cg.flush(); cg.run = 0x10 | op; }
	}
	int qb = x + 1, tb = y + 1;
	if(qb){
		if(1u == (cg.run & 0xf)) cg.run += 0x10u * (uint32_t)qb; else { cg.flush(); cg.run = (0x10u * (uint32_t)qb) | 1u; }
		ins += qb; qb = 0;
	}
	if((type == 0 || type == 2) && tb){
		if(2u == (cg.run & 0xf)) cg.run += 0x10u * (uint32_t)tb; else { cg.flush(); cg.run = (0x10u * (uint32_t)tb) | 2u; }
		del += tb; tb = 0;
	}
	cg.flush();
	int score;
	if(type == 1) score = smin + te - tb;
	else score = smin;
	rs[0] = score; rs[1] = qb; rs[2] = qe; rs[3] = tb; rs[4] = te;
	rs[5] = mat; rs[6] = mis; rs[7] = ins; rs[8] = del; rs[9] = mat + mis + ins + del;
	emit_dense(cg, a.dense, a.dense_off, a.dense_total, a.ncigar, pair);
	a.status[pair] = err | cg.err;
	#undef TP
	#undef RBEG
	#undef QB
}

} // namespace bsb200
