// kmer_edit.cuh -- the k-mer guided edit alignment (replaces kmer_striped_seqedit_pairwise, bsalign.h:1209-1536; callers: main.c:196
// `bsalign edit -m kmer -k ksz`, bspoa.h:2089 band placement of long reads) for BATCHES of pairs, all of it on the device.
//
// The reference sorts the canonical k-mers of both sequences, keeps those seen exactly twice (once per sequence, same strand), sorts
// them by query offset, chains them, filters the chain by diagonal and aligns the gaps between the anchors with the edit DP.  Here:
//   * persistent warps fetch GROUPS of up to 32 pairs from a counter; every warp owns one scratch slot.  The warp-wide phases run pair
//     after pair, the serial phases on one lane per pair (group = 1 when a batch has fewer pairs than the GPU has warps).
//   * unique shared k-mers come from a hash table in the warp's slot instead of two sorts: every lane rolls over its own stretch of
//     positions and inserts (atomicCAS on the key; a per-sequence 64-bit word takes occurrences << 32 | offset|strand as a fire-and-forget
//     add), then the query positions are scanned IN ORDER and a ballot compacts
//     the hits - which is the reference's list after its second sort (query offsets are distinct).
//   * the chain (patience tails with the reference's own predecessor rule), the diagonal filter (mean / median / 3x rule) and the
//     coverage tests are sequential per pair: one lane.
//   * the gaps between the anchors are independent edit alignments: the gaps of all pairs of a group form one work list that is
//     spread over the lanes - a Myers/Hyyro bit-vector DP of
//     (gap length / 64) words per row whose trace lives in the lane's own scratch (one-word gaps up to 127 rows) or in a pool
//     (anything larger; a pair that finds the pool empty is flagged and run again by the host once the pool is free).
//   * one lane stitches the gap cigars around the anchor matches with the reference's push order (bsalign.h:1461-1531) and the warp
//     copies the result into the dense cigar arena.
// A pair without usable anchors is flagged kStFallback: the host runs those through the plain global edit kernel (bsalign.h:1440).
#pragma once
#include "common.cuh"
#include <stdint.h>

namespace bsb200 {

constexpr int kKmWarps = 4;                    // warps per CTA
constexpr uint32_t kKmSmallRows = 128;         // trace rows (incl. the init row) of a one-word gap that fit the lane's own scratch
constexpr int kStFallback = 0x40000000;        // internal: no usable anchors, the pair takes the plain global edit
constexpr int kStPool = 0x08000000;            // internal: the pool for large gap traces was exhausted, run the pair again
constexpr uint32_t kKmNone = 0xFFFFFFFFu, kKmOn = 0x80000000u;

struct KmerArgs {
	const uint8_t *seqs;
	const uint64_t *qoff, *toff;
	const uint32_t *qlen, *tlen;
	const uint32_t *order;           // pairs of this launch (nullptr: 0 .. npairs-1)
	uint32_t npairs, ksz;
	uint32_t group;                  // pairs a warp works on at a time (1..32)
	uint32_t smem_keys;              // hash-table keys per warp kept in shared memory (0: the keys live in the warp's slot)
	unsigned int *next;              // pair counter
	uint8_t *scratch; uint64_t warp_bytes;
	uint64_t off_kq, off_lane, off_glist, off_gmeta, off_pairs, pair_bytes;        // a warp's slot: hash table, query k-mers, lane scratch, gap work list, then `group` pair blocks
	uint64_t off_hq, off_ht, off_tails, off_pred, off_seg, off_stage, off_out;     // byte offsets inside a pair block
	uint8_t *pool; unsigned long long *pool_used; uint64_t pool_bytes;
	int32_t *results, *status;
	uint32_t *ncigar, *dense; uint64_t *dense_off; unsigned long long *dense_total;
};

// scratch a warp needs for `group` pairs up to (maxq, maxt) at a time; fills the offsets of `a`
__host__ inline uint64_t kmer_warp_bytes(uint32_t maxq, uint32_t maxt, uint32_t group, KmerArgs *a){
	uint64_t H = 64; while(2 * H < 3 * ((uint64_t)maxq + maxt)) H <<= 1;
	const uint64_t mh = (uint64_t)(maxq < maxt ? maxq : maxt) + 2;
	auto up = [](uint64_t x){ return (x + 127) / 128 * 128; };
	uint64_t o = up(H * 20);
	const uint64_t off_kq = o; o += up(((uint64_t)maxq + 2) * 4);
	const uint64_t off_lane = o; o += (uint64_t)32 * 2 * kKmSmallRows * 8;
	const uint64_t off_gmeta = o; o += 3 * 32 * 4 + 128;
	const uint64_t off_glist = o; o += up((mh + 1) * 4 * group);
	const uint64_t off_pairs = o;
	uint64_t q = 0;
	const uint64_t off_hq = q; q += up(mh * 4);
	const uint64_t off_ht = q; q += up(mh * 4);
	const uint64_t off_tails = q; q += up(mh * 4);
	const uint64_t off_pred = q; q += up(mh * 4);
	const uint64_t off_seg = q; q += up((mh + 1) * 32);
	const uint64_t off_stage = q; q += up(((uint64_t)maxq + maxt + 8) * 4);
	const uint64_t off_out = q; q += up(((uint64_t)maxq + maxt + 8) * 4);
	o += q * group;
	if(a){
		a->off_kq = off_kq; a->off_lane = off_lane; a->off_glist = off_glist; a->off_gmeta = off_gmeta; a->off_pairs = off_pairs; a->pair_bytes = q; a->group = group;
		a->off_hq = off_hq; a->off_ht = off_ht; a->off_tails = off_tails; a->off_pred = off_pred; a->off_seg = off_seg; a->off_stage = off_stage; a->off_out = off_out;
		a->warp_bytes = o;
	}
	return o;
}

__device__ __forceinline__ uint64_t km_lowmask(uint32_t n){ return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

__device__ __forceinline__ uint32_t km_slot(uint32_t km, uint32_t hbits){ return (km * 2654435761u) >> (32 - hbits); }

// k-th smallest of a[0..n) (the reference's quick_median_array, sort.h:268-310, returns the element of rank n/2)
__device__ inline int km_select(int *a, int n, int k){
	int lo = 0, hi = n - 1;
	while(lo < hi){
		const int piv = a[(lo + hi) >> 1];
		int i = lo, j = hi;
		do {
			while(a[i] < piv) i++;
			while(a[j] > piv) j--;
			if(i <= j){ int t_ = a[i]; a[i] = a[j]; a[j] = t_; i++; j--; }
		} while(i <= j);
		if(k <= j) hi = j; else if(k >= i) lo = i; else break;
	}
	return a[k];
}

// One gap between two anchors: striped_seqedit_pairwise (bsalign.h:1046) in GLOBAL (type 0) or EXTEND (type 2) mode with the band
// disabled (bandwidth 0 = the whole query, so the band never moves), forward sweep and backtrace.  q[x] = qp[x * qd], t[y] = tp[y * td]
// (qd = td = -1 walks the reversed prefixes of the stretch before the first anchor, bsalign.h:1490-1494).  Scratch words are
// scr[k * sd].  The cigar words are written in BACKTRACE order (end to start).  rec: qe, te, mat, mis, ins, del, score, words.
__device__ inline int km_gap(const uint8_t *qp, int qd, uint32_t sq, const uint8_t *tp, int td, uint32_t st, int type,
		uint64_t *scr, uint32_t sd, uint32_t *cig, uint32_t cigcap, int32_t *rec){
	const uint32_t W = (sq + 63) / 64;
	uint64_t ql1 = 0, qh1 = 0;   // the query planes of a one-word gap
	#define KS(k) scr[(size_t)(k) * sd]
	#define KT(row, w, plane) KS(2 * (size_t)W + (((size_t)(row) * W + (w)) * 2 + (plane)))   // plane 0 = minus, 1 = plus
	for(uint32_t w=0;w<W;w++){
		uint64_t lo = 0, hi = 0;
		const uint32_t x0 = w * 64, n = sq - x0 < 64 ? sq - x0 : 64;
		for(uint32_t k=0;k<n;k++){ const uint32_t c = qp[(int64_t)(x0 + k) * qd]; lo |= (uint64_t)(c & 1) << k; hi |= (uint64_t)((c >> 1) & 1) << k; }
		KS(w) = lo; KS(W + w) = hi;
		KT(0, w, 0) = 0ull; KT(0, w, 1) = ~0ull;
		ql1 = lo; qh1 = hi;
	}
	int sbeg = 0, smin = 0x7FFFFFFF, rx = (int)sq - 1, ry = (int)st - 1;
	uint64_t pv1 = ~0ull, mv1 = 0ull;   // the row of a one-word gap stays in registers
	for(uint32_t i=0;i<st;i++){
		const uint32_t tb = tp[(int64_t)i * td];
		const uint64_t TL = (tb & 1) ? ~0ull : 0ull, TH = (tb & 2) ? ~0ull : 0ull;
		uint64_t carry = 0, phin = type == 1 ? 0ull : 1ull, mhin = 0;   // OVERLAP (type 1, the long-pair kernel below): H(-1, y) = 0
		int rowsum = 0;
		if(type != 1) sbeg++;
		for(uint32_t w=0;w<W;w++){
			const uint64_t ql = W == 1 ? ql1 : KS(w), qh = W == 1 ? qh1 : KS(W + w);
			const uint64_t valid = km_lowmask(sq - 64 * w);
			const uint64_t Eq = ~(ql ^ TL) & ~(qh ^ TH) & valid;
			uint64_t pv = W == 1 ? pv1 : KT(i, w, 1), mv = W == 1 ? mv1 : KT(i, w, 0);
			const uint64_t Xv = Eq | mv;
			const uint64_t a0 = Eq & pv;
			const uint64_t sum = a0 + pv;
			uint64_t c1 = sum < a0;
			const uint64_t sum2 = sum + carry;
			c1 |= (sum2 < sum);
			carry = c1;
			const uint64_t Xh = (sum2 ^ pv) | Eq;
			uint64_t Ph = mv | ~(Xh | pv);
			uint64_t Mh = pv & Xh;
			const uint64_t pho = Ph >> 63, mho = Mh >> 63;
			Ph = (Ph << 1) | phin; Mh = (Mh << 1) | mhin;
			phin = pho; mhin = mho;
			pv = Mh | ~(Xv | Ph);
			mv = Ph & Xv;
			if(W == 1){ pv1 = pv; mv1 = mv; }
			KT(i + 1, w, 0) = mv; KT(i + 1, w, 1) = pv;
			rowsum += __popcll(pv & valid) - __popcll(mv & valid);
		}
		if(type != 0){ const int srow = sbeg + rowsum; if(srow < smin){ smin = srow; rx = (int)sq - 1; ry = (int)i; } }   // bsalign.h:1124-1139
		else if(i + 1 == st) smin = sbeg + rowsum;
	}
	if(type == 2 && W == 1){ // arg-min over the last row, one word: bit-lane j is band position j (the general form is below)
		int run = sbeg, best = sbeg; uint32_t pmin = 0;
		for(uint32_t blk4=0;blk4<4;blk4++){
			int sc = 0; uint32_t st_ = 0;
			for(uint32_t l=0;l<16;l++){
				const uint32_t j = blk4 * 16 + l;
				const int u = (int)((pv1 >> j) & 1) - (int)((mv1 >> j) & 1);
				const int c = run + (u < 0 ? u : 0); run += u;
				if(l == 0 || sc > c){ sc = c; st_ = l; }
			}
			if(sc >= best) continue;
			best = sc; pmin = blk4 * 16 + st_;
		}
		if(best < smin){ smin = best; rx = (int)pmin; ry = (int)st - 1; }
	} else if(type == 2){ // arg-min over the last row in the reference's lane / chunk order (bsalign.h:813-963)
		int sb = sbeg, best = sbeg; uint32_t pmin = 0;
		uint32_t cw = W == 1 ? 0u : 0xffffffffu; uint64_t pw_ = pv1, mw_ = mv1;   // the word of the last row the scan is in
		for(uint32_t blk4=0;blk4<4;blk4++){
			int sc = 0; uint32_t st_ = 0, stp = 0;
			int run = sb;
			for(uint32_t l=0;l<16;l++){
				const uint32_t j = blk4 * 16 + l;   // bit-lane j covers band positions [j*W, (j+1)*W)
				int hh = 0, mm = 0; uint32_t ppos = 0;
				for(uint32_t ib=0;ib<W;ib+=124){
					const uint32_t ie = ib + 124 < W ? ib + 124 : W;
					int h = 0, m = 0; uint32_t pz = 0;
					for(uint32_t k=ib;k<ie;k++){
						const uint32_t p = j * W + k;
						if((p >> 6) != cw){ cw = p >> 6; pw_ = KT(st, cw, 1); mw_ = KT(st, cw, 0); }
						h += (int)((pw_ >> (p & 63)) & 1) - (int)((mw_ >> (p & 63)) & 1);
						if(m > h){ m = h; pz = k - ib; }
					}
					const int d = hh + m;
					if(mm > d){ mm = d; ppos = pz + ib; }
					hh += h;
				}
				const int c = run + mm; run += hh;
				if(l == 0 || sc > c){ sc = c; st_ = l; stp = ppos; }
			}
			sb = run;
			if(sc >= best) continue;
			best = sc; pmin = (blk4 * 16 + st_) * W + stp;
		}
		if(best < smin){ smin = best; rx = (int)pmin; ry = (int)st - 1; }
	}
	// backtrace (bsalign.h:965-1044)
	uint32_t n = 0, run = 0; int err = 0;
	auto flush = [&](){ if(run){ if(n < cigcap) cig[n] = run; else err |= 4; n++; } run = 0; };
	int x = rx, y = ry, mat = 0, mis = 0, ins = 0, del = 0;
	rec[0] = x + 1; rec[1] = y + 1;
	int64_t guard = 0;
	while(x >= 0 && y >= 0){
		if(++guard > 4 * ((int64_t)sq + st) + 64){ err |= 2; break; }
		uint32_t op;
		if(qp[(int64_t)x * qd] == tp[(int64_t)y * td]){ mat++; op = 0; x--; y--; }
		else {
			const uint64_t w1p = KT(y + 1, x >> 6, 1), w1m = KT(y + 1, x >> 6, 0), w0p = KT(y, x >> 6, 1), w0m = KT(y, x >> 6, 0);
			const int u_here = (int)((w1p >> (x & 63)) & 1) - (int)((w1m >> (x & 63)) & 1);
			if(u_here == 1){ ins++; op = 1; x--; }
			else {
				const int u_up = (int)((w0p >> (x & 63)) & 1) - (int)((w0m >> (x & 63)) & 1);
				if(u_up == -1){ del++; op = 2; y--; }
				else { mis++; op = 0; x--; y--; }
			}
		}
		if(op == (run & 0xf)) run += 0x10;
		else { flush(); run = 0x10 | op; }
	}
	const int qb = x + 1, tb_ = y + 1;
	if(qb){
		if(1u == (run & 0xf)) run += 0x10u * (uint32_t)qb; else { flush(); run = (0x10u * (uint32_t)qb) | 1u; }
		ins += qb;
	}
	if(tb_ && type != 1){ // GLOBAL and EXTEND both pad the target (bsalign.h:1029)
		if(2u == (run & 0xf)) run += 0x10u * (uint32_t)tb_; else { flush(); run = (0x10u * (uint32_t)tb_) | 2u; }
		del += tb_;
	}
	flush();
	rec[2] = mat; rec[3] = mis; rec[4] = ins; rec[5] = del; rec[6] = smin; rec[7] = (int32_t)n;
	#undef KS
	#undef KT
	return err;
}

// per-pair arrays inside a warp's slot (pair j of the group the warp is working on)
struct KmView { uint32_t *hq, *ht, *tails, *pred; int32_t *seg; uint32_t *stage, *out; };
__device__ __forceinline__ KmView km_view(const KmerArgs &a, uint8_t *ws, uint32_t j){
	uint8_t *pb = ws + a.off_pairs + (uint64_t)j * a.pair_bytes;
	KmView v;
	v.hq = (uint32_t*)(pb + a.off_hq); v.ht = (uint32_t*)(pb + a.off_ht); v.tails = (uint32_t*)(pb + a.off_tails); v.pred = (uint32_t*)(pb + a.off_pred);
	v.seg = (int32_t*)(pb + a.off_seg); v.stage = (uint32_t*)(pb + a.off_stage); v.out = (uint32_t*)(pb + a.off_out);
	return v;
}

__device__ __forceinline__ uint32_t km_cmin(uint32_t qlen, uint32_t tlen, uint32_t ksz){   // bsalign.h:1221-1222, the reference's double arithmetic
	const uint32_t mn = qlen < tlen ? qlen : tlen;
	const uint32_t c = (uint32_t)__dadd_rn(__dmul_rn((double)mn, 0.05), 1.0);
	return c > 2 * ksz ? 2 * ksz : c;
}

// no alignment from this kernel: all-zero result, status says why (one lane)
__device__ __forceinline__ void km_leave(const KmerArgs &a, uint32_t pair, int st){
	int32_t *rs = a.results + (size_t)pair * 10;
	for(int k=0;k<10;k++) rs[k] = 0;
	a.status[pair] = st; a.ncigar[pair] = 0; a.dense_off[pair] = 0;
}

// Phase A (whole warp): unique shared canonical k-mers of one pair (bsalign.h:1230-1276) into v.hq / v.ht in query order.
// Returns the number of hits, or -1 when the pair is finished already (empty input, or no chance of anchors: flagged for the fallback).
__device__ inline int km_hits(const KmerArgs &a, const uint32_t pair, uint8_t *ws, const KmView &v, const uint32_t lane){
	const uint32_t FULL = 0xffffffffu;
	const uint32_t qlen = a.qlen[pair], tlen = a.tlen[pair], ksz = a.ksz;
	const uint8_t *qs = a.seqs + a.qoff[pair], *ts = a.seqs + a.toff[pair];
	if(qlen == 0 || tlen == 0){ if(lane == 0) km_leave(a, pair, 0); return -1; }
	const uint32_t nq = qlen >= ksz ? qlen - ksz + 1 : 0, nt = tlen >= ksz ? tlen - ksz + 1 : 0;
	if(nq == 0 || nt == 0){ if(lane == 0) km_leave(a, pair, kStFallback); return -1; }
	uint32_t hbits = 6; while(2 * (1ull << hbits) < 3 * ((uint64_t)nq + nt)) hbits++;   // load <= 2/3
	const uint32_t H = 1u << hbits, hm = H - 1;
	// table: keys[H] (all ones = empty), then per sequence one 64-bit word per slot: occurrences << 32 | sum of (offset << 1 | strand).
	// Only the key needs an atomic with a result (probing); the occurrence words are fire-and-forget adds.
	// The keys of short pairs sit in SHARED memory (a.smem_keys entries per warp: the key atomics and the probe loads are the
	// dependent accesses of this phase); the occurrence words stay in the warp's slot.
	extern __shared__ uint32_t km_smem[];
	uint32_t *keys = a.smem_keys ? km_smem + (size_t)(threadIdx.x >> 5) * a.smem_keys : (uint32_t*)ws;
	unsigned long long *qv = (unsigned long long*)((uint32_t*)ws + H), *tv = qv + H;
	{
		const uint4 f = make_uint4(kKmNone, kKmNone, kKmNone, kKmNone), o = make_uint4(0, 0, 0, 0);
		uint4 *zk = (uint4*)keys, *zv = (uint4*)qv;
		for(uint32_t i=lane;i<H/4;i+=32) zk[i] = f;
		for(uint32_t i=lane;i<H;i+=32) zv[i] = o;
	}
	__syncwarp();
	// every lane rolls over its own stretch of k-mer positions of BOTH sequences (one byte per position and sequence; the two key
	// atomics of an iteration are in flight together); the query's k-mers are kept for the scan below
	uint32_t *kq = (uint32_t*)(ws + a.off_kq);
	const uint32_t kmk = 0xFFFFFFFFu >> ((16 - ksz) << 1), sft = (ksz - 1) << 1;
	{
		const uint32_t perq = (nq + 31) / 32, pert = (nt + 31) / 32;
		uint32_t pq = lane * perq, pt = lane * pert;
		const uint32_t eq = pq + perq < nq ? pq + perq : nq, et = pt + pert < nt ? pt + pert : nt;
		uint32_t fq = 0, rq = 0, ft = 0, rt = 0;
		if(pq < eq) for(uint32_t i=0;i+1<ksz;i++){ const uint32_t c = qs[pq + i] & 3u; fq = (fq << 2) | c; rq = (rq >> 2) | ((3u - c) << sft); }
		if(pt < et) for(uint32_t i=0;i+1<ksz;i++){ const uint32_t c = ts[pt + i] & 3u; ft = (ft << 2) | c; rt = (rt >> 2) | ((3u - c) << sft); }
		while(pq < eq || pt < et){
			const bool vq = pq < eq, vt = pt < et;
			uint32_t kmq = 0, kmt = 0, dq = 0, dt = 0, hq_ = 0, ht_ = 0, oq = 0, ot = 0;
			if(vq){ const uint32_t c = qs[pq + ksz - 1] & 3u; fq = ((fq << 2) | c) & kmk; rq = (rq >> 2) | ((3u - c) << sft); dq = rq < fq; kmq = dq ? rq : fq; hq_ = km_slot(kmq, hbits); }
			if(vt){ const uint32_t c = ts[pt + ksz - 1] & 3u; ft = ((ft << 2) | c) & kmk; rt = (rt >> 2) | ((3u - c) << sft); dt = rt < ft; kmt = dt ? rt : ft; ht_ = km_slot(kmt, hbits); }
			if(vq) oq = atomicCAS(&keys[hq_], kKmNone, kmq);
			if(vt) ot = atomicCAS(&keys[ht_], kKmNone, kmt);
			if(vq){
				while(!(oq == kKmNone || oq == kmq)){ hq_ = (hq_ + 1) & hm; oq = atomicCAS(&keys[hq_], kKmNone, kmq); }
				atomicAdd(&qv[hq_], (1ull << 32) | ((pq << 1) | dq));
				kq[pq] = (kmq << 1) | dq;
				pq++;
			}
			if(vt){
				while(!(ot == kKmNone || ot == kmt)){ ht_ = (ht_ + 1) & hm; ot = atomicCAS(&keys[ht_], kKmNone, kmt); }
				atomicAdd(&tv[ht_], (1ull << 32) | ((pt << 1) | dt));
				pt++;
			}
		}
	}
	__syncwarp();
	uint32_t nh = 0;
	for(uint32_t base=0;base<nq;base+=32){
		const uint32_t p = base + lane;
		bool hit = false; uint32_t tt = 0;
		if(p < nq){
			const uint32_t km = kq[p] >> 1;
			uint32_t h = km_slot(km, hbits);
			while(keys[h] != km) h = (h + 1) & hm;
			const unsigned long long x = qv[h], w = tv[h];
			hit = (x >> 32) == 1 && (w >> 32) == 1 && ((x ^ w) & 1u) == 0;
			// the reference's scan ends on a zeroed sentinel, so a LAST group of k-mer value 0 is never looked at: only possible when it is the only one
			if(km == 0 && nq == 1 && nt == 1) hit = false;
			tt = (uint32_t)w >> 1;
		}
		const uint32_t mask = __ballot_sync(FULL, hit);
		if(hit){ const uint32_t r = nh + __popc(mask & ((1u << lane) - 1u)); v.hq[r] = p; v.ht[r] = tt; }
		nh += __popc(mask);
	}
	__syncwarp();
	if(nh * ksz < km_cmin(qlen, tlen, ksz)){ if(lane == 0) km_leave(a, pair, kStFallback); return -1; }
	return (int)nh;
}

// Phase B (ONE lane per pair): chain, diagonal filter, coverage (bsalign.h:1277-1424); the anchors are compacted in place.  Returns
// their number, 0 = the pair takes the fallback.
__device__ inline uint32_t km_chain(const KmView &v, const uint32_t nh, const uint32_t ksz, const uint32_t cmin){
	uint32_t *hq = v.hq, *ht = v.ht, *tails = v.tails, *pred = v.pred;
	uint32_t len = 1, b, e, m;
	tails[0] = 0; pred[0] = kKmNone;
	for(uint32_t i=1;i<nh;i++){
		const uint32_t ti = ht[i];
		if(ti > ht[tails[len - 1]]){ pred[i] = tails[len - 1]; tails[len++] = i; }
		else if(ti <= ht[tails[0]]){ pred[i] = kKmNone; tails[0] = i; }
		else {
			b = 0; e = len;
			while(b < e){
				m = b + ((e - b) >> 1);
				const uint32_t tm = ht[tails[m]];
				if(ti > tm) b = m + 1; else if(ti < tm) e = m; else { b = m; break; }
			}
			pred[i] = pred[tails[b - 1]];   // the reference's rule: the predecessor of the tail before it
			tails[b] = i;
		}
	}
	b = 0; e = 0xFFFFFFFFu;
	for(m=tails[len-1];m!=kKmNone;m=pred[m]){
		hq[m] |= kKmOn;
		if(ht[m] + ksz <= e) b += ksz; else b += e - ht[m];
		e = ht[m];
	}
	if(b < cmin) return 0;
	int *dl = (int*)tails;
	for(;;){ // drop anchors whose diagonal is far from the mean (bsalign.h:1347-1394)
		int tot = 0; uint32_t cnt = 0, drop = 0;
		for(uint32_t i=0;i<nh;i++) if(hq[i] & kKmOn){ const int d = (int)(hq[i] & ~kKmOn) - (int)ht[i]; tot += d; dl[cnt++] = d; }
		if(cnt * ksz < cmin) break;
		const int mean = tot / (int)cnt;
		const int median = km_select(dl, (int)cnt, (int)(cnt / 2));
		int var = (median > mean ? median - mean : mean - median) * 3;
		if(var < 50) var = 50;
		for(uint32_t i=0;i<nh;i++) if(hq[i] & kKmOn){
			int d = (int)(hq[i] & ~kKmOn) - (int)ht[i] - mean;
			if(d < 0) d = -d;
			if(d > var){ hq[i] &= ~kKmOn; drop++; }
		}
		if(drop == 0) break;
	}
	uint32_t kept = 0; m = 0; e = 0;
	for(uint32_t i=0;i<nh;i++) if(hq[i] & kKmOn){
		const uint32_t t_ = ht[i];
		if(t_ >= e + ksz) m += ksz; else m += t_ + ksz - e;
		e = t_ + ksz;
		hq[kept] = hq[i] & ~kKmOn; ht[kept] = t_; kept++;
	}
	return m >= cmin ? kept : 0;
}

// Phase C1 (ONE lane per pair): the gaps of a pair that need a DP (bsalign.h:1461-1531) are appended to the warp's work list as
// (pair block << 27 | gap index); adjacent anchors have none, a gap with one empty side gets the all-zero record of bsalign.h:1051-1054.
// pass 0 counts, pass 1 writes from `at`.
__device__ inline uint32_t km_gap_list(const KmerArgs &a, const uint32_t pair, const KmView &v, const uint32_t kmap, const uint32_t blk, uint32_t *list, uint32_t at, const int pass){
	const uint32_t qlen = a.qlen[pair], tlen = a.tlen[pair], kh = a.ksz / 2;
	const uint32_t *hq = v.hq, *ht = v.ht;
	uint32_t c = 0;
	for(uint32_t i=0;i<=kmap;i++){
		const uint32_t qb = i ? hq[i - 1] + kh + 1 : 0, tb = i ? ht[i - 1] + kh + 1 : 0;
		const uint32_t qe = i < kmap ? hq[i] + kh : qlen, te = i < kmap ? ht[i] + kh : tlen;
		if(qb == qe && tb == te) continue;
		if(qe == qb || te == tb){ if(pass){ int32_t *rec = v.seg + (size_t)i * 8; for(int k=0;k<8;k++) rec[k] = 0; } continue; }
		if(pass) list[at + c] = (blk << 27) | i;
		c++;
	}
	return c;
}

// Phase C2 (any lane): one entry of the work list
__device__ inline int km_gap_run(const KmerArgs &a, const uint32_t pair, uint8_t *ws, const KmView &v, const uint32_t kmap, const uint32_t i, const uint32_t lane){
	const uint32_t qlen = a.qlen[pair], tlen = a.tlen[pair], kh = a.ksz / 2;
	const uint8_t *qs = a.seqs + a.qoff[pair], *ts = a.seqs + a.toff[pair];
	const uint32_t *hq = v.hq, *ht = v.ht;
	const uint32_t qb = i ? hq[i - 1] + kh + 1 : 0, tb = i ? ht[i - 1] + kh + 1 : 0;
	const uint32_t qe = i < kmap ? hq[i] + kh : qlen, te = i < kmap ? ht[i] + kh : tlen;
	int32_t *rec = v.seg + (size_t)i * 8;
	const uint32_t sq = qe - qb, st = te - tb;
	const uint32_t W = (sq + 63) / 64;
	const uint64_t words = 2 * (uint64_t)W * ((uint64_t)st + 2);
	uint64_t *scr; uint32_t sd;
	if(W == 1 && st + 1 <= kKmSmallRows - 1){ scr = (uint64_t*)(ws + a.off_lane) + lane; sd = 32; }
	else {
		const unsigned long long bytes = (words * 8 + 15) & ~15ull;
		const unsigned long long off = atomicAdd(a.pool_used, bytes);
		if(off + bytes > a.pool_bytes){ rec[7] = 0; return kStPool; }
		scr = (uint64_t*)(a.pool + off); sd = 1;
	}
	uint32_t *cg = v.stage + qb + tb;
	if(i == 0) return km_gap(qs + qe - 1, -1, sq, ts + te - 1, -1, st, 2, scr, sd, cg, sq + st, rec);
	return km_gap(qs + qb, 1, sq, ts + tb, 1, st, i == kmap ? 2 : 0, scr, sd, cg, sq + st, rec);
}

// Phase D (ONE lane per pair): stitch the gap cigars around the anchor matches with the reference's push order; result, status, count.
__device__ inline uint32_t km_stitch(const KmerArgs &a, const uint32_t pair, const KmView &v, const uint32_t kmap, int gerr){
	const uint32_t qlen = a.qlen[pair], tlen = a.tlen[pair], kh = a.ksz / 2;
	const uint32_t *hq = v.hq, *ht = v.ht;
	uint32_t *out = v.out;
	const uint32_t outcap = qlen + tlen + 2;
	uint32_t n = 0, ml = 0;
	int R[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	auto raw = [&](uint32_t w){ if(n < outcap) out[n] = w; else gerr |= 4; n++; };
	for(uint32_t i=0;i<=kmap;i++){
		const uint32_t qb = i ? hq[i - 1] + kh + 1 : 0, tb = i ? ht[i - 1] + kh + 1 : 0;
		const uint32_t qe = i < kmap ? hq[i] + kh : qlen, te = i < kmap ? ht[i] + kh : tlen;
		if(i < kmap) ml++;
		if(qb == qe && tb == te) continue;
		const int32_t *rec = v.seg + (size_t)i * 8;
		const uint32_t *cg = v.stage + qb + tb;
		const uint32_t nw = (uint32_t)rec[7];
		if(i == 0){ // the reference pushes M, appends the gap's cigar and reverses everything (bsalign.h:1476-1499)
			for(uint32_t k=0;k<nw;k++) raw(cg[k]);
			raw(ml << 4);
			R[5] += (int)ml; R[9] += (int)ml; ml = 0;
			R[1] = (int)qe - rec[0]; R[3] = (int)te - rec[1]; R[2] = (int)qe; R[4] = (int)te;
		} else {
			if(ml){
				if(n && n <= outcap && (out[n - 1] & 0xf) == 0) out[n - 1] += ml << 4; else raw(ml << 4);
				R[5] += (int)ml; R[9] += (int)ml; ml = 0;
			}
			for(uint32_t k=0;k<nw;k++) raw(cg[nw - 1 - k]);
			R[2] = (int)qb + rec[0]; R[4] = (int)tb + rec[1];
		}
		R[5] += rec[2]; R[6] += rec[3]; R[7] += rec[4]; R[8] += rec[5]; R[9] += rec[2] + rec[3] + rec[4] + rec[5]; R[0] += rec[6];
	}
	int32_t *rs = a.results + (size_t)pair * 10;
	for(int k=0;k<10;k++) rs[k] = R[k];
	if(n > outcap) n = outcap;
	a.ncigar[pair] = n;
	a.status[pair] = gerr & 7;
	a.dense_off[pair] = 0;
	return n;
}

// A warp works on a.group pairs at a time: the warp-wide phases (A, C and the copy) run pair after pair, the serial phases (B, D) on
// one lane per pair.  group = 1 (few long pairs: more warps than pairs) leaves the serial phases to lane 0.
__global__ void __launch_bounds__(kKmWarps * 32, 8) kmer_edit_kernel(const KmerArgs a){
	const uint32_t FULL = 0xffffffffu;
	const uint32_t lane = threadIdx.x & 31, G = a.group;
	uint8_t *ws = a.scratch + (uint64_t)(blockIdx.x * kKmWarps + (threadIdx.x >> 5)) * a.warp_bytes;
	for(;;){
		uint32_t idx0 = 0;
		if(lane == 0) idx0 = atomicAdd(a.next, G);
		idx0 = __shfl_sync(FULL, idx0, 0);
		if(idx0 >= a.npairs) break;
		const uint32_t cnt = a.npairs - idx0 < G ? a.npairs - idx0 : G;
		uint32_t my_pair = 0, my_kmap = 0, my_n = 0; int my_nh = -1, my_gerr = 0;
		for(uint32_t j=0;j<cnt;j++){
			const uint32_t pair = a.order ? a.order[idx0 + j] : idx0 + j;
			const int nh = km_hits(a, pair, ws, km_view(a, ws, j), lane);
			if(lane == j){ my_pair = pair; my_nh = nh; }
			__syncwarp();
		}
		if(my_nh >= 0){
			my_kmap = km_chain(km_view(a, ws, lane), (uint32_t)my_nh, a.ksz, km_cmin(a.qlen[my_pair], a.tlen[my_pair], a.ksz));
			if(my_kmap == 0) km_leave(a, my_pair, kStFallback);
		}
		__syncwarp();
		// the gaps of all pairs of the group form one work list, spread evenly over the lanes
		uint32_t *glist = (uint32_t*)(ws + a.off_glist), *gmeta = (uint32_t*)(ws + a.off_gmeta);   // gmeta: [pair | kmap | flags] x 32
		const KmView mine = km_view(a, ws, lane);
		const uint32_t c = my_kmap ? km_gap_list(a, my_pair, mine, my_kmap, lane, glist, 0, 0) : 0;
		uint32_t incl = c;
		for(int o=1;o<32;o<<=1){ const uint32_t t_ = __shfl_up_sync(FULL, incl, o); if(lane >= (uint32_t)o) incl += t_; }
		const uint32_t total = __shfl_sync(FULL, incl, 31);
		if(my_kmap) km_gap_list(a, my_pair, mine, my_kmap, lane, glist, incl - c, 1);
		gmeta[lane] = my_pair; gmeta[32 + lane] = my_kmap; gmeta[64 + lane] = 0;
		__syncwarp();
		for(uint32_t e=lane;e<total;e+=32){
			const uint32_t w = glist[e], blk = w >> 27, gi = w & 0x7FFFFFFu;
			const int g = km_gap_run(a, gmeta[blk], ws, km_view(a, ws, blk), gmeta[32 + blk], gi, lane);
			if(g) atomicOr(&gmeta[64 + blk], (uint32_t)g);
		}
		__syncwarp();
		my_gerr = (int)gmeta[64 + lane];
		if(my_kmap){
			if(my_gerr & kStPool) km_leave(a, my_pair, kStPool);
			else my_n = km_stitch(a, my_pair, km_view(a, ws, lane), my_kmap, my_gerr);
		}
		__syncwarp();
		for(uint32_t j=0;j<cnt;j++){
			const uint32_t n = __shfl_sync(FULL, my_n, j), pair = __shfl_sync(FULL, my_pair, j);
			if(n == 0) continue;
			unsigned long long off = 0;
			if(lane == 0){ off = atomicAdd(a.dense_total, (unsigned long long)n); a.dense_off[pair] = off; }
			off = __shfl_sync(FULL, off, 0);
			const uint32_t *out = km_view(a, ws, j).out;
			for(uint32_t k=lane;k<n;k+=32) a.dense[off + k] = out[k];
		}
		__syncwarp();
	}
}

// ---- unbanded edit alignment of pairs too long for edit_kernel (band > 16384 cells, or a query whose bit-planes outgrow shared memory) ----
// One thread per pair runs the general form of km_gap above on scratch in HBM: as slow as one CPU core, but striped_seqedit_pairwise
// (bsalign.h:1046) has no length limit and a drop-in must not have one either.  Modes GLOBAL (whole query as the band), OVERLAP, EXTEND.
struct EditLongArgs {
	const uint8_t *seqs; const uint64_t *qoff, *toff; const uint32_t *qlen, *tlen;
	const uint32_t *pairs; const uint64_t *scr_off; uint32_t npairs;
	uint8_t *scratch; int mode;
	int32_t *results, *status; uint32_t *cigars; const uint64_t *cig_off;
	uint32_t *dense; uint64_t *dense_off; unsigned long long *dense_total; uint32_t *ncigar;
};

__global__ void __launch_bounds__(32) edit_long_kernel(const EditLongArgs a){
	if(threadIdx.x || blockIdx.x >= a.npairs) return;
	const uint32_t pair = a.pairs[blockIdx.x];
	const uint32_t qlen = a.qlen[pair], tlen = a.tlen[pair];
	const int type = a.mode & 3;
	int32_t rec[8];
	CigarSink cg;
	cg.buf = a.cigars ? a.cigars + a.cig_off[pair] : nullptr;
	cg.cap = a.cigars ? (uint32_t)(a.cig_off[pair + 1] - a.cig_off[pair]) : 0;
	cg.n = 0; cg.run = 0; cg.err = 0;
	uint32_t dummy = 0;
	const int err = km_gap(a.seqs + a.qoff[pair], 1, qlen, a.seqs + a.toff[pair], 1, tlen, type, (uint64_t*)(a.scratch + a.scr_off[blockIdx.x]), 1,
		cg.buf ? cg.buf : &dummy, cg.cap, rec);
	cg.n = (uint32_t)rec[7];
	const int qe = rec[0], te = rec[1], mat = rec[2], mis = rec[3], ins = rec[4], del = rec[5];
	const int tb = type == 1 ? te - (mat + mis + del) : 0;   // GLOBAL / EXTEND pad the target down to 0
	int32_t *rs = a.results + (size_t)pair * 10;
	rs[0] = type == 1 ? rec[6] + te - tb : rec[6];
	rs[1] = 0; rs[2] = qe; rs[3] = tb; rs[4] = te; rs[5] = mat; rs[6] = mis; rs[7] = ins; rs[8] = del; rs[9] = mat + mis + ins + del;
	if(cg.buf && cg.n > cg.cap) cg.err |= 4;
	emit_dense(cg, a.dense, a.dense_off, a.dense_total, a.ncigar, pair);
	a.status[pair] = ((cg.buf ? err : (err & ~4)) | cg.err) & 7;
}

// results of the pairs that took the plain global edit (a sub-batch) go back into the batch's arrays; their cigar words are appended
// to the dense arena
__global__ void __launch_bounds__(256) kmer_merge_kernel(const uint32_t *idx, uint32_t m, const int32_t *sres, const int32_t *sstat, const uint32_t *sncg,
		const uint32_t *sdense, const uint64_t *sprefix, int32_t *results, int32_t *status, uint32_t *ncigar, uint32_t *dense, uint64_t *dense_off,
		unsigned long long *dense_total){
	const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if(f >= m) return;
	const uint32_t pair = idx[f], n = sncg[f];
	if(lane < 10) results[(size_t)pair * 10 + lane] = sres[(size_t)f * 10 + lane];
	unsigned long long off = 0;
	if(lane == 0){ status[pair] = sstat[f]; ncigar[pair] = n; off = atomicAdd(dense_total, (unsigned long long)n); dense_off[pair] = off; }
	off = __shfl_sync(0xffffffffu, off, 0);
	const uint32_t *src = sdense + sprefix[f];
	for(uint32_t k=lane;k<n;k+=32) dense[off + k] = src[k];
}

} // namespace bsb200
