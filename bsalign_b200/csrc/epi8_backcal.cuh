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
	int anch;                    // rows carry sub-lane anchors (written by the ANCH forward instantiations)
	int ubias;                   // 128 when the forward kernel stored u + 128 (FAST instantiations), else 0
	int split;                   // sub-blocks per lane of the wavefront kernel that wrote the skewed pairs (epi8_wave.cuh)
	uint32_t stride;             // one walk every `stride` threads (power of two <= 32): small batches spread their walks over more warps
	int8_t mtx[16];
	int8_t go1, ge1, go2, ge2;
};

struct TraceView {
	const uint8_t *tr; const int32_t *meta; uint32_t bw, W, IB, RS, AOFF; int tlen, ubias, anch;
	// skew = 1: the pair was written by the wavefront kernel (epi8_wave.cuh): lane j's row y sits in image slot y + 1 + j, the end
	// anchor ub[k] (k >= 1) of row y in record y + k of `meta` (16 ints per record), ub[0] in `ub0`, the band never moved, the two
	// steps of a word are step-major (epi8_cell_offset_w) and e is stored + 128
	// With `split` sub-blocks per lane the stage of (lane j, step i) is split * j + i / Wb, and that many slots further down.
	int skew, split; uint32_t Wb; const int32_t *ub0;
	// p / W without a division (the walk does it twice per step): one multiply with ceil(2^32 / W) and a correction; exact for p < 16 W, W < 16384
	uint32_t Wmagic;
	__device__ __forceinline__ uint32_t div_w(uint32_t p) const { if(W == 1) return p; uint32_t j = __umulhi(p, Wmagic); return j * W > p ? j - 1 : j; }
	__device__ __forceinline__ int beg(int row) const { return skew ? 0 : meta[(size_t)kMetaInts * (row + 1) + 17]; }
	__device__ __forceinline__ int ub(int row, int j) const {
		if(skew) return j ? meta[(size_t)16 * (row + split * j) + j - 1] : ub0[row + 1];
		return meta[(size_t)kMetaInts * (row + 1) + j];
	}
	// first byte of the image slot that holds step i of lane j of a row
	__device__ __forceinline__ const uint8_t *slot(int row, uint32_t j, uint32_t i) const { return tr + (size_t)RS * (row + 1 + (skew ? (int)(split * j + i / Wb) : 0)); }
	// array arr (0 u, 1 e, 2 q) of the cell at band position p of a row
	__device__ __forceinline__ int cell(int row, int arr, uint32_t p) const {
		uint32_t j = div_w(p), i = p - j * W;
		uint8_t b = skew ? slot(row, j, i)[epi8_trace_offset_w(j, i) + 16 * arr] : slot(row, j, i)[(size_t)arr * IB + epi8_cell_offset(j, i)];
		return (int)(int8_t)(((arr == 0 && ubias) || (arr == 1 && skew)) ? (uint8_t)(b ^ 0x80) : b);
	}
	// H(col,row) = anchor of the lane + its u cells up to col (bsalign.h:3187-3202); sets err when the lookup leaves the band
	__device__ int score(int row, int col, int &err) const { return score_at(row, (row >= -1 && row < tlen) ? beg(row) : 0, col, err); }
	// ---- split lookup: `begin` only computes addresses and issues the loads (anchor + up to 8 chunks), `finish` sums.
	// Between the two the caller issues its other loads, so that one walk step costs a single memory round trip.
	struct Pending { const uint8_t *r; int4 v[8]; int ub; uint32_t n, cs; int mk; };
	// a word holds two steps of the lane: mask of the first one
	__device__ __forceinline__ int first_step(int mk) const { return skew ? (mk & 0x00ff00ff) : (mk & 0x0000ffff); }
	__device__ __forceinline__ void begin(Pending &p, bool need, int row, int rbeg, int col, int &err) const {
		int64_t pos = (int64_t)col - rbeg;
		bool ok = need && row >= -1 && row < tlen && pos >= 0 && pos < (int64_t)bw;
		if(need && !ok) err |= 1;
		uint32_t up = ok ? (uint32_t)pos : 0u, j = div_w(up);
		int rw = ok ? row : -1;
		const uint32_t i = up - j * W, g = anch ? i / kAnchorSteps : 0u;   // the sum starts at the sub-lane anchor before step 32g
		p.n = ok ? i - g * kAnchorSteps + 1 : 0u;
		const uint8_t *sl = slot(rw, j, i);
		// (the sub-lane anchor before step 32g was written by the stage of step 32g - 1)
		p.ub = ok ? (g ? *(const int32_t*)(slot(rw, j, g * kAnchorSteps - 1) + AOFF + ((g - 1) * 16 + j) * 4) : ub(rw, j)) : kScoreMin;
		p.r = skew ? sl + (size_t)(j >> 1) * 32 + (size_t)g * (kAnchorSteps / 8) * 256 : sl + (size_t)(j >> 1) * 16 + (size_t)g * (kAnchorSteps / 8) * 128;
		p.cs = skew ? 256u : 128u;   // bytes from one chunk of the lane pair to the next
		p.mk = skew ? ((j & 1) ? 0x01010000 : 0x00000101) : ((j & 1) ? 0x01000100 : 0x00010001);
		const uint32_t nch = (p.n + 7) >> 3;
		#pragma unroll
		for(int k=0;k<8;k++) p.v[k] = (uint32_t)k < nch ? *(const int4*)(p.r + (size_t)p.cs * k) : make_int4(0, 0, 0, 0);
	}
	// (branch-free: a word holds two steps of the lane; the mask of a word is the lane's, the first step's, or 0)
	__device__ __forceinline__ int word_mask(int left, int mk) const { return left >= 2 ? mk : (left == 1 ? first_step(mk) : 0); }
	__device__ __forceinline__ int finish(const Pending &p) const {
		int s = p.ub;
		const int xm = ubias ? (int)0x80808080 : 0; // u + 128 as unsigned byte -> two's complement
		#pragma unroll
		for(int k=0;k<8;k++){
			const int left = (int)p.n - 8 * k;
			if(left <= 0) break;   // (one test per chunk; the words of a chunk are branch-free)
			s = __dp4a(p.v[k].x ^ xm, word_mask(left, p.mk), s);
			s = __dp4a(p.v[k].y ^ xm, word_mask(left - 2, p.mk), s);
			s = __dp4a(p.v[k].z ^ xm, word_mask(left - 4, p.mk), s);
			s = __dp4a(p.v[k].w ^ xm, word_mask(left - 6, p.mk), s);
		}
		const uint32_t nch = (p.n + 7) >> 3;
		for(uint32_t c0=8;c0<nch;c0+=8){ // lanes longer than 64 steps without anchors: further rounds
			int4 v[8];
			#pragma unroll
			for(int k=0;k<8;k++) v[k] = c0 + k < nch ? *(const int4*)(p.r + (size_t)p.cs * (c0 + k)) : make_int4(0, 0, 0, 0);
			#pragma unroll
			for(int k=0;k<8;k++){
				const int left = (int)p.n - 8 * (int)(c0 + k);
				s = __dp4a(v[k].x ^ xm, word_mask(left, p.mk), s);
				s = __dp4a(v[k].y ^ xm, word_mask(left - 2, p.mk), s);
				s = __dp4a(v[k].z ^ xm, word_mask(left - 4, p.mk), s);
				s = __dp4a(v[k].w ^ xm, word_mask(left - 6, p.mk), s);
			}
		}
		return s;
	}
	// same with the row's band offset supplied by the caller (the walk keeps the offsets of rows tb, tb-1, tb-2 in registers)
	__device__ int score_at(int row, int rbeg, int col, int &err) const {
		Pending p;
		begin(p, true, row, rbeg, col, err);
		return finish(p);
	}
};

__global__ void __launch_bounds__(128) epi8_backcal_kernel(const Epi8BtArgs a){
	uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	// The 32 walks of a warp run in lock step: every iteration waits for the slowest of 32 scattered loads and executes the union of
	// the branches taken.  A batch that cannot fill the GPU with warps anyway puts one walk every `stride` threads.
	if(idx & (a.stride - 1)) return;
	idx /= a.stride;
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
	tv.W = tv.bw / kLanes; tv.Wmagic = tv.W > 1 ? (uint32_t)((0x100000000ull + tv.W - 1) / tv.W) : 0xffffffffu; tv.IB = epi8_image_bytes(tv.W); tv.RS = a.anch ? epi8_row_bytes(tv.W, pw) : tv.IB * (pw + 1); tv.AOFF = tv.IB * (pw + 1); tv.anch = a.anch; tv.tlen = tlen; tv.ubias = a.ubias;
	tv.tr = a.trace + a.trace_off[pair];
	int err = a.status[pair];
	tv.skew = (err & kStSkew) ? 1 : 0;
	err &= ~(kStSkew | kStRedo);
	tv.split = a.split > 0 ? a.split : 1;
	tv.Wb = tv.split > 1 ? ((tv.W + tv.split - 1) / tv.split + kStageAlign - 1) / kStageAlign * kStageAlign : (tv.W + 7) / 8 * 8;
	const uint32_t nslot = (uint32_t)tlen + 1 + (tv.skew ? (uint32_t)(kLanes * tv.split - 1) : 0u);
	tv.meta = (const int32_t*)(tv.tr + (size_t)tv.RS * nslot);
	tv.ub0 = tv.meta + (size_t)16 * nslot;
	const int bw = (int)tv.bw;
	CigarSink cg;
	cg.buf = a.cigars ? a.cigars + a.cig_off[pair] : nullptr;
	cg.cap = a.cigars ? (uint32_t)(a.cig_off[pair + 1] - a.cig_off[pair]) : 0;
	cg.n = 0; cg.run = 0; cg.err = 0;
	int qb = rs[2], tb = rs[4];
	int mat = 0, mis = 0, ins = 0, del = 0, aln = 0;
	int Hcur, Hprev = 0, pend = 0, prior = 0;
	int64_t guard = 0; const int64_t guard_max = 16 * ((int64_t)qlen + tlen) + 64;
	const int qe = qb + 1, te = tb + 1;
	// band offsets of rows tb, tb-1, tb-2: the one for tb-2 is requested a full step before it is needed, which
	// takes the anchor record out of the dependent-load chain of a step
	int b0 = tv.beg(tb), b1 = tb >= 0 ? tv.beg(tb - 1) : 0, b2 = tb >= 1 ? tv.beg(tb - 2) : 0;
	#define ROW_UP() { tb--; b0 = b1; b1 = b2; b2 = tb >= 1 ? tv.beg(tb - 2) : 0; }
	Hcur = tv.score_at(tb, b0, qb, err);
	// The walk of bsalign.h:3726-3821 as a state machine with ONE score lookup per iteration: the 32 pairs of a warp
	// take different branches all the time, and with one shared lookup site per iteration a warp pays one memory round
	// trip per iteration instead of one per branch taken by any of its threads.
	enum { kStep = 0, kDrun = 1, kIsearch = 2 };
	int state = kStep, sz = 0;
	while(true){
		if(++guard > guard_max){ err |= 2; break; }
		// ---- 1. what this iteration needs -----------------------------------------------------------------
		int lrow = 0, lbeg = 0, lcol = 0, x = 0;
		bool need = false, cellok = false;
		if(state == kStep){
			if(qb < 0 || tb < 0) break;
			x = qb - b1;
			if(qb == b1){
				if(qb){ Hprev = tv.ub(tb - 1, 0); prior = 0; } // left edge of the previous row: only a deletion can follow (:3761-3765)
				else if(mode == 1 || tb == 0) Hprev = 0;
				else if(pw < 2) Hprev = go1 + ge1 * tb;
				else Hprev = max(go1 + ge1 * tb, go2 + ge2 * tb);
			} else if(x <= bw){ need = true; lrow = tb - 1; lbeg = b1; lcol = qb - 1; }
			cellok = (x >= 0 && x < bw);
		} else if(state == kDrun){
			need = true; lrow = tb; lbeg = b0; lcol = qb;
		} else {
			if(sz + b0 > qb){ err |= 2; break; } // no insertion length fits: the reference repeats this step forever (:3798-3814)
			need = true; lrow = tb; lbeg = b0; lcol = qb - sz;
		}
		// ---- 2. issue every load of this step (cell above, query/target bases, lookup), then consume -------------
		const int crow = (state == kStep) ? tb - 1 : -1;                    // row of the cell above (valid memory in any state)
		const uint32_t cx = cellok ? (uint32_t)x : 0u;
		const uint32_t cj = tv.div_w(cx), coff = tv.skew ? epi8_trace_offset_w(cj, cx - cj * tv.W) : epi8_cell_offset(cj, cx - cj * tv.W);
		const uint8_t *cp = tv.slot(crow, cj, cx - cj * tv.W) + coff;
		const uint32_t ru = cp[0], re = pw >= 1 ? cp[tv.skew ? 16 : tv.IB] : 0u, rq = pw == 2 ? cp[2 * (size_t)tv.IB] : 0u;
		const uint32_t qbase = qs[qb >= 0 ? qb : 0], tbase = ts[tb >= 0 ? tb : 0];
		TraceView::Pending pd;
		tv.begin(pd, need, lrow, lbeg, lcol, err);
		int val = tv.finish(pd);
		int u = 0, e = (int)(int8_t)(go1 + ge1), q = 0;
		if(cellok){ u = (int)(int8_t)(a.ubias ? (ru ^ 0x80u) : ru); if(pw >= 1) e = (int)(int8_t)(tv.skew ? (re ^ 0x80u) : re); if(pw == 2) q = (int)(int8_t)rq; }
		else { u = 0; e = 0; q = 0; }
		// ---- 3. act ------------------------------------------------------------------------------------------------
		if(state == kDrun){
			const int len = pend >> 4;
			const int cost = ((pend & 0xf) == 2) ? go1 + len * ge1 : go2 + len * ge2;
			if(val + cost == Hcur){
				cg.push(2, len);
				del += len; aln += len;
				Hcur = val; pend = 0; state = kStep;
			} else { pend += 1 << 4; ROW_UP(); }
		} else if(state == kIsearch){
			int tt = go1 + sz * ge1;
			if(pw == 2) tt = max(tt, go2 + sz * ge2);
			if(val + tt == Hcur){
				cg.push(1, sz);
				Hcur = val; qb -= sz; ins += sz; aln += sz;
				state = kStep;
			} else sz++;
		} else {
			if(need) Hprev = val;
			const int s = a.mtx[qbase * 4 + tbase];
			const int h = Hcur - Hprev;
			int bt;
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
				if(qbase == tbase) mat++; else mis++;
				qb--; aln++;
				ROW_UP();
				cg.push(0, 1);
				Hcur = Hprev;
			} else if(bt == 1){
				if(qb <= 0){
					cg.push(1, 1);
					Hcur = Hprev;
					qb--; ins++; aln++;
				} else { state = kIsearch; sz = 1; }
			} else {
				pend = (1 << 4) | bt;
				ROW_UP();
				state = kDrun;
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
