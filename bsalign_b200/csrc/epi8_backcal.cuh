// epi8_backcal.cuh -- traceback of the 8-bit banded DP by re-derivation (replaces bsalign.h:3667-3852).
// One thread per pair walks the HBM trace written by epi8_forward_kernel; every score lookup is the lane
// anchor plus a masked dp4a sum over at most W bytes.
#pragma once
#include "common.cuh"

namespace bsb200 {

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
	const uint8_t *tr; const int32_t *meta; uint32_t bw, W, IB, RS; int tlen;
	__device__ __forceinline__ int beg(int row) const { return meta[(size_t)kMetaInts * (row + 1) + 17]; }
	__device__ __forceinline__ int ub(int row, int j) const { return meta[(size_t)kMetaInts * (row + 1) + j]; }
	// array arr (0 u, 1 e, 2 q) of the cell at band position p of a row
	__device__ __forceinline__ int cell(int row, int arr, uint32_t p) const {
		uint32_t j = p / W, i = p - j * W;
		return (int)(int8_t)tr[(size_t)RS * (row + 1) + (size_t)arr * IB + epi8_cell_offset(j, i)];
	}
	// H(col,row) = anchor of the lane + its u cells up to col (bsalign.h:3187-3202); sets err when the lookup leaves the band
	__device__ int score(int row, int col, int &err) const { return score_at(row, (row >= -1 && row < tlen) ? beg(row) : 0, col, err); }
	// same with the row's band offset supplied by the caller (the walk keeps the offsets of rows tb, tb-1, tb-2 in registers)
	__device__ int score_at(int row, int rbeg, int col, int &err) const {
		if(row < -1 || row >= tlen){ err |= 1; return kScoreMin; }
		int64_t pos = (int64_t)col - rbeg;
		if(pos < 0 || pos >= (int64_t)bw){ err |= 1; return kScoreMin; }
		uint32_t j = (uint32_t)pos / W, n = (uint32_t)pos - j * W + 1;
		int s = ub(row, j);
		// the lane's cells are every other byte of its thread's 16 bytes per chunk: dp4a with a 0/1 mask adds two per word
		const uint8_t *r = tr + (size_t)RS * (row + 1) + (size_t)(j >> 1) * 16;
		const int mk = (j & 1) ? 0x01000100 : 0x00010001;
		// all chunk loads of a round are issued before the first sum so that a lookup costs one memory round trip
		// per 8 chunks (64 steps of the lane) instead of one per chunk
		const uint32_t nch = (n + 7) >> 3;
		for(uint32_t c0=0;c0<nch;c0+=8){
			int4 v[8];
			#pragma unroll
			for(int k=0;k<8;k++) if(c0 + k < nch) v[k] = *(const int4*)(r + 128 * (c0 + k));
			#pragma unroll
			for(int k=0;k<8;k++){
				if(c0 + k >= nch) break;
				const uint32_t left = n - 8 * (c0 + k); // entries of this chunk that count (>= 1)
				const int w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
				#pragma unroll
				for(int q=0;q<4;q++){
					if(left >= 2u * q + 2) s = __dp4a(w[q], mk, s);
					else if(left == 2u * q + 1) s = __dp4a(w[q], mk & 0x0000ffff, s);
				}
			}
		}
		return s;
	}
};

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
	tv.W = tv.bw / kLanes; tv.IB = epi8_image_bytes(tv.W); tv.RS = tv.IB * (pw + 1); tv.tlen = tlen;
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
	// band offsets of rows tb, tb-1, tb-2: the one for tb-2 is requested a full step before it is needed, which
	// takes the anchor record out of the dependent-load chain of a step
	int b0 = tv.beg(tb), b1 = tb >= 0 ? tv.beg(tb - 1) : 0, b2 = tb >= 1 ? tv.beg(tb - 2) : 0;
	#define ROW_UP() { tb--; b0 = b1; b1 = b2; b2 = tb >= 1 ? tv.beg(tb - 2) : 0; }
	Hcur = tv.score_at(tb, b0, qb, err);
	while(true){
		if(++guard > guard_max){ err |= 2; break; }
		if((pend & 0xf) == 2 || (pend & 0xf) == 4){
			int len = pend >> 4;
			Hprev = tv.score_at(tb, b0, qb, err);
			int cost = ((pend & 0xf) == 2) ? go1 + len * ge1 : go2 + len * ge2;
			if(Hprev + cost == Hcur){
				cg.push(2, len);
				del += len; aln += len;
				Hcur = Hprev; pend = 0;
			} else { pend += 1 << 4; ROW_UP(); continue; }
		}
		if(qb < 0 || tb < 0) break;
		const int pbeg = b1;
		if(qb == pbeg){
			if(qb){ Hprev = tv.ub(tb - 1, 0); prior = 0; }
			else if(mode == 1 || tb == 0) Hprev = 0;
			else if(pw < 2) Hprev = go1 + ge1 * tb;
			else Hprev = max(go1 + ge1 * tb, go2 + ge2 * tb);
		} else if(qb - pbeg <= bw){
			Hprev = tv.score_at(tb - 1, b1, qb - 1, err);
		}
		{
			const int x = qb - pbeg;
			const int s = a.mtx[qs[qb] * 4 + ts[tb]];
			const int h = Hcur - Hprev;
			int bt;
			if(x > bw) bt = 1;
			else if(x == bw) bt = (h == s) ? 0 : 1;
			else if(prior && h == s) bt = 0;    // the common step: no need to look at the cell above
			else {
				int u = 0, e = 0, q = 0;
				if(x >= 0 && x < bw){
					u = tv.cell(tb - 1, 0, (uint32_t)x);
					e = pw >= 1 ? tv.cell(tb - 1, 1, (uint32_t)x) : (int)(int8_t)(go1 + ge1);
					q = pw == 2 ? tv.cell(tb - 1, 2, (uint32_t)x) : 0;
				}
				if(h == u + e) bt = 2;
				else if(pw == 2 && h == u + q) bt = 4;
				else if(!prior && h == s) bt = 0;
				else bt = 1;
			}
			prior = 1;
			if(bt == 0){
				if(qs[qb] == ts[tb]) mat++; else mis++;
				qb--; aln++;
				ROW_UP();
				cg.push(0, 1);
				Hcur = Hprev;
			} else if(bt == 1){
				if(qb <= 0){
					cg.push(1, 1);
					Hcur = Hprev;
					qb--; ins++; aln++;
				} else {
					const int cbeg = b0;
					for(int sz=1;sz+cbeg<=qb;sz++){
						int tt = go1 + sz * ge1;
						if(pw == 2) tt = max(tt, go2 + sz * ge2);
						int Hl = tv.score_at(tb, b0, qb - sz, err);
						if(Hl + tt == Hcur){
							cg.push(1, sz);
							Hcur = Hl; qb -= sz; ins += sz; aln += sz;
							break;
						}
					}
				}
			} else {
				pend = (1 << 4) | bt;
				ROW_UP();
				continue;
			}
		}
	}
	#undef ROW_UP
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
