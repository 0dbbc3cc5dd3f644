/*
 * oracle/bsalign_oracle.c -- TEST INFRASTRUCTURE ONLY. Never linked into, imported by or executed from
 * the product path (bsalign_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * A plain scalar C restatement of the reference's banded striped DP hot path, written in LINEAR band
 * coordinates (band position p = 0..bw-1; the reference's SSE lane j owns the "running block"
 * p in [j*W, (j+1)*W), W = bw/16).  Every int8 operation saturates exactly where the SSE code
 * saturates (adds/subs_epi8), so the output is bit-identical to the reference's SSE4.2 build.
 *
 * Parity status: PINNED.  tests/test_oracle.py (epi8, edit) and tests/test_poa.py (POA sweep, walk) compare every function here against the
 * unmodified reference compiled into oracle/_ref/libbsref.so (see oracle/ref_harness.c), and
 * tests/golden/ holds vectors generated from that build (tests/golden/make_golden.py) plus the README
 * example (README.md:36-42).
 *
 * Citations are into /root/reference/bsalign.h unless stated otherwise.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define BSO_LANES 16                 /* WORDSIZE of the SSE4.2 build, bsalign.h:142 */
#define BSO_EPI8_MIN (-63)           /* SEQALIGN_SCORE_EPI8_MIN, bsalign.h:56 */
#define BSO_EPI8_MAX (63)            /* SEQALIGN_SCORE_EPI8_MAX, bsalign.h:57 */
#define BSO_SCORE_MIN (-536870911)   /* SEQALIGN_SCORE_MIN = -(INT32_MAX >> 2), bsalign.h:58 */
#define BSO_MODE_GLOBAL 0
#define BSO_MODE_OVERLAP 1
#define BSO_MODE_EXTEND 2

#define BSO_ERR_RANGE 1   /* a score lookup left the band (the reference would read out of bounds) */
#define BSO_ERR_LOOP  2   /* traceback found no consistent predecessor (the reference would spin forever) */
#define BSO_ERR_CIGCAP 4  /* cigar buffer too small */
#define BSO_ERR_REFBUG 8  /* input reaches a reference code path that corrupts its own scratch memory (edit row_movx, bsalign.h:704-713) */

static inline int8_t sat8(int v){ return (int8_t)(v > 127 ? 127 : (v < -128 ? -128 : v)); }
static inline int8_t adds8(int8_t a, int8_t b){ return sat8((int)a + (int)b); }
static inline int8_t subs8(int8_t a, int8_t b){ return sat8((int)a - (int)b); }
static inline int8_t max8(int8_t a, int8_t b){ return a > b ? a : b; }

/* bsalign.h:2084-2092 */
int bso_epi8_piecewise(int8_t go1, int8_t ge1, int8_t go2, int8_t ge2, int bw){
	if(go2 < go1 && ge2 > ge1 && go2 + ge2 < go1 + ge1 && (go1 - go2) / (ge1 - ge2) < bw) return 2;
	return go1 ? 1 : 0;
}

typedef struct {
	int8_t *u, *e, *q;  /* linear band order */
	int32_t *ub;        /* 17 ints: ub[0] = H just left of u[0]; ub[j] = H at band pos j*W-1 */
} bso_row_t;

typedef struct {
	uint32_t qlen, tlen, bw, W;
	int pw, mode;
	const uint8_t *qseq, *tseq;
	const int8_t *mtx;
	int8_t go1, ge1, go2, ge2;
	int8_t smax, smin;
	int8_t hpc;     /* POA only: bonus added where the next query base differs (set_query_prof_hpc, :2194-2221); 0 otherwise */
	/* trace: row y lives at index y+1 (row -1 = init row, bsalign.h:3922) */
	int8_t *U, *E, *Q;
	int32_t *UB;    /* (tlen+1) * 17 */
	int32_t *begs;  /* tlen+1, begs[0] is row -1 */
	int err;
} bso_epi8_t;

static inline bso_row_t bso_trace_row(bso_epi8_t *a, int64_t y){
	bso_row_t r;
	r.u = a->U + (y + 1) * (int64_t)a->bw;
	r.e = a->E ? a->E + (y + 1) * (int64_t)a->bw : NULL;
	r.q = a->Q ? a->Q + (y + 1) * (int64_t)a->bw : NULL;
	r.ub = a->UB + (y + 1) * 17;
	return r;
}

/* bsalign.h:2094-2140: row -1 */
static void bso_row_init(bso_epi8_t *a, bso_row_t r, int8_t max_nt, int8_t min_nt){
	uint32_t p, j, W = a->W, bw = a->bw;
	int32_t blk[BSO_LANES], s;
	int two = (a->go2 < a->go1 && a->ge2 > a->ge1 && a->go2 + a->ge2 < a->go1 + a->ge1 && (a->go1 - a->go2) / (a->ge1 - a->ge2) < (int)bw);
	if(a->mode == BSO_MODE_GLOBAL || a->mode == BSO_MODE_EXTEND){
		int8_t ext = two ? a->ge2 : a->ge1;
		for(p=0;p<bw;p++) r.u[p] = ext;
		for(j=0;j<BSO_LANES;j++) blk[j] = ext * (int)W;
		r.u[0] = (int8_t)(a->go1 + a->ge1 + min_nt - max_nt);
		blk[0] += r.u[0] - ext;
		if(two){
			uint32_t xp = (uint32_t)((a->go2 - a->go1) / (a->ge1 - a->ge2));
			for(p=1;p<xp&&p<bw;p++){ /* the reference has no p<bw guard: xp >= bw writes past the row */
				r.u[p] = a->ge1;
				blk[p / W] += a->ge1 - a->ge2;
			}
		}
		s = max_nt - min_nt;
		for(j=0;j<BSO_LANES;j++){ r.ub[j] = s; s += blk[j]; }
		r.ub[BSO_LANES] = s;
	} else {
		memset(r.u, 0, bw);
		memset(r.ub, 0, 17 * sizeof(int32_t));
	}
	if(two){
		memset(r.e, BSO_EPI8_MIN, bw);
		memset(r.q, BSO_EPI8_MIN, bw);
	} else if(a->go1){
		memset(r.e, BSO_EPI8_MIN, bw);
	}
}

/* bsalign.h:3187-3197: absolute H at band position pos of a row */
static int bso_getscore(bso_epi8_t *a, bso_row_t r, int64_t pos){
	uint32_t W = a->W, j, i, k;
	int s;
	if(pos < 0 || pos >= (int64_t)a->bw){ a->err |= BSO_ERR_RANGE; return BSO_SCORE_MIN; }
	j = pos / W; i = pos % W;
	s = r.ub[j];
	for(k=0;k<=i;k++) s += r.u[j * W + k];
	return s;
}

/* bsalign.h:2244-2392: realign the previous row to a band that starts m cells further right */
static void bso_row_shift(bso_epi8_t *a, bso_row_t src, bso_row_t dst, uint32_t m, int8_t nt_max, int8_t nt_min){
	uint32_t W = a->W, bw = a->bw, p, j, cyc, mov;
	int32_t sums[BSO_LANES];
	int pw = a->pw;
	if(m >= bw){ /* :2253-2259 */
		memset(dst.u, 0, bw);
		if(pw) memset(dst.e, 0, bw);
		if(pw == 2) memset(dst.q, 0, bw);
		for(j=0;j<=BSO_LANES;j++) dst.ub[j] = BSO_SCORE_MIN;
		return;
	}
	if(m == 0){ /* :2260-2269 */
		memcpy(dst.u, src.u, bw);
		if(pw) memcpy(dst.e, src.e, bw);
		if(pw == 2) memcpy(dst.q, src.q, bw);
		memcpy(dst.ub, src.ub, 17 * sizeof(int32_t));
		return;
	}
	cyc = m / W; mov = m % W;
	/* the byte shuffles of :2271-2345 amount to dst[p] = src[p + m], zero filled past the old band */
	for(p=0;p<bw;p++){
		int in = (p + m < bw);
		dst.u[p] = in ? src.u[p + m] : 0;
		if(pw) dst.e[p] = in ? src.e[p + m] : 0;
		if(pw == 2) dst.q[p] = in ? src.q[p + m] : 0;
	}
	/* :2310-2331, :2350-2355: block anchors; sums[j] = old ub[j] + first `mov` cells of old block j */
	for(j=0;j<BSO_LANES;j++){
		sums[j] = src.ub[j];
		for(p=0;p<mov;p++) sums[j] += src.u[j * W + p];
	}
	for(j=0;j+cyc<BSO_LANES;j++) dst.ub[j] = sums[j + cyc];
	for(j=BSO_LANES-cyc;j<=BSO_LANES;j++) dst.ub[j] = src.ub[BSO_LANES];
	/* :2357-2389: synthesize the overhang (cells the old band did not cover) */
	{
		uint32_t i0 = bw - m, d;
		int c;
		if(pw == 2){
			d = (uint32_t)((a->go1 - a->go2) / (a->ge2 - a->ge1));
			c = (nt_min < a->go2 + a->ge2 ? nt_min : a->go2 + a->ge2) - 1 - nt_max + (a->go2 + a->ge2);
		} else {
			d = bw + 1;
			c = (nt_min < a->go1 + a->ge1 ? nt_min : a->go1 + a->ge1) - 1 - nt_max + (a->go1 + a->ge1);
		}
		dst.u[i0] = (int8_t)c;
		/* c accumulates the overhang values; every block end crossed adds the running total to its anchor */
		for(p=i0+1;p<=bw;p++){
			if(p % W == 0){ /* crossed the end of block p/W - 1 */
				dst.ub[p / W] += c;
			}
			if(p == bw) break;
			{
				int8_t g = (p < i0 + d) ? a->ge1 : a->ge2;
				dst.u[p] = g;
				c += g;
			}
		}
	}
}

/* score of aligning query position x against target base tb: the query-profile entry (:2166-2191) */
static inline int8_t bso_prof(bso_epi8_t *a, uint32_t x, uint8_t tb){
	int c;
	if(x >= a->qlen) return BSO_EPI8_MIN;
	c = a->mtx[a->qseq[x] * 4 + tb];
	if(a->hpc && x + 1 < a->qlen && a->qseq[x] != a->qseq[x + 1]) c += a->hpc; /* :2204-2206 */
	return (int8_t)c;
}

/* :2639-2652: carry the block-exit F of lane j-1 into lane j, extended across whole blocks */
static void bso_fpen(bso_epi8_t *a, const int8_t fend[BSO_LANES], const int32_t *ub, int8_t ge, int8_t fin[BSO_LANES]){
	int j, s, t;
	fin[0] = BSO_EPI8_MIN;
	for(j=1;j<BSO_LANES;j++) fin[j] = fend[j - 1];
	t = (int)a->W * ge;
	s = t + fin[0] - (ub[1] - ub[0]);
	for(j=1;j<BSO_LANES;j++){
		if(fin[j] < s) fin[j] = (int8_t)s;
		s = t + fin[j] - (ub[j + 1] - ub[j]);
	}
}

/*
 * One DP row (bsalign.h:2727-2793 linear, :2885-2960 affine, :3084-3179 two-piece, tail :2618-2636).
 * prev = previous row already shifted to this row's band; out = new row.  Returns nothing; out.ub[0]
 * is H(rbeg, y).
 */
static void bso_row_cal(bso_epi8_t *a, uint32_t rbeg, uint8_t tb, bso_row_t prev, bso_row_t out, int rh){
	uint32_t W = a->W, j, i, p;
	int pw = a->pw;
	int8_t GE = a->ge1, GOE = (int8_t)(a->go1 + a->ge1), GP = a->ge2, GQP = (int8_t)(a->go2 + a->ge2);
	int8_t GOQ = subs8(GOE, GQP);
	int8_t fend[BSO_LANES], gend[BSO_LANES], fin[BSO_LANES], gin[BSO_LANES], vt[BSO_LANES];
	int h0, t0;
	/* cell 0 of the band: the diagonal predecessor H(rbeg-1, y-1) is outside the stored row (:2899-2907) */
	h0 = (rh - prev.ub[0]) + bso_prof(a, rbeg, tb);
	if(pw == 0) t0 = prev.u[0] + GE;
	else if(pw == 1) t0 = prev.u[0] + prev.e[0];
	else t0 = prev.u[0] + (prev.e[0] > prev.q[0] ? prev.e[0] : prev.q[0]);
	if(h0 >= t0){ if(h0 > BSO_EPI8_MAX) h0 = BSO_EPI8_MAX; }
	else h0 = BSO_EPI8_MIN;
	/* pass 1: F (and G) leaving every running block when nothing enters it */
	for(j=0;j<BSO_LANES;j++){
		int8_t f = BSO_EPI8_MIN, g = BSO_EPI8_MIN, h, u, e, q;
		for(i=0;i<W;i++){
			p = j * W + i;
			h = (p == 0) ? (int8_t)h0 : bso_prof(a, rbeg + p, tb);
			u = prev.u[p];
			if(pw == 0){
				e = adds8(u, GE);
				h = max8(e, h); h = max8(f, h);
				f = subs8(adds8(h, GE), u);
			} else if(pw == 1){
				e = adds8(prev.e[p], u);
				h = max8(e, h); h = max8(f, h);
				f = adds8(f, GE); h = adds8(h, GOE); f = max8(f, h); f = subs8(f, u);
			} else {
				e = adds8(prev.e[p], u); q = adds8(prev.q[p], u);
				h = max8(e, h); h = max8(q, h); h = max8(f, h); h = max8(g, h);
				f = adds8(f, GE); h = adds8(h, GOE); f = max8(f, h); f = subs8(f, u);
				g = adds8(g, GP); h = subs8(h, GOQ); g = max8(g, h); g = subs8(g, u);
			}
		}
		fend[j] = f; gend[j] = g;
	}
	bso_fpen(a, fend, prev.ub, GE, fin);
	if(pw == 2) bso_fpen(a, gend, prev.ub, GP, gin);
	/* pass 2: the real row */
	for(j=0;j<BSO_LANES;j++){
		int8_t f = fin[j], g = (pw == 2) ? gin[j] : 0, h = 0, u = 0, e, q, v = 0, z;
		for(i=0;i<W;i++){
			p = j * W + i;
			z = (p == 0) ? (int8_t)h0 : bso_prof(a, rbeg + p, tb);
			u = prev.u[p];
			if(pw == 0){
				e = adds8(u, GE);
				h = max8(e, z); h = max8(f, h);
				out.u[p] = subs8(h, v);
				v = subs8(h, u);
				f = subs8(adds8(h, GE), u);
			} else if(pw == 1){
				e = adds8(prev.e[p], u);
				h = max8(e, z); h = max8(f, h);
				out.u[p] = subs8(h, v);
				v = subs8(h, u);
				e = adds8(e, GE); e = subs8(e, h); e = max8(e, GOE);
				out.e[p] = e;
				f = adds8(f, GE); h = adds8(h, GOE); f = max8(f, h); f = subs8(f, u);
			} else {
				e = adds8(prev.e[p], u);
				h = max8(e, z);
				q = adds8(prev.q[p], u);
				h = max8(q, h); h = max8(f, h); h = max8(g, h);
				out.u[p] = subs8(h, v);
				v = subs8(h, u);
				e = adds8(e, GE); e = subs8(e, h); e = max8(e, GOE);
				out.e[p] = e;
				q = adds8(q, GP); q = subs8(q, h); q = max8(q, GQP);
				out.q[p] = q;
				f = adds8(f, GE); h = adds8(h, GOE); f = max8(f, h); f = subs8(f, u);
				g = adds8(g, GP); h = subs8(h, GOQ); g = max8(g, h); g = subs8(g, u);
			}
		}
		/* the SSE code leaves h biased by the gap-open constant after the loop and removes it (:2958, :3177) */
		if(pw == 1) h = subs8(h, GOE);
		else if(pw == 2) h = subs8(h, GQP);
		vt[j] = subs8(h, u); /* V = H(x,y) - H(x,y-1) of the block's last cell (:2622) */
	}
	/* tail (:2618-2636): new anchors; first cell of every block still lacks the V of its left neighbour */
	for(j=1;j<=BSO_LANES;j++) out.ub[j] = prev.ub[j] + vt[j - 1];
	for(j=1;j<BSO_LANES;j++) out.u[j * W] = subs8(out.u[j * W], vt[j - 1]);
	out.ub[0] = prev.ub[0] + out.u[0];
	out.u[0] = 0;
}

/* bsalign.h:3213-3291: arg-max of a row with the SSE reduction's tie-break order */
static uint32_t bso_row_max(bso_epi8_t *a, bso_row_t r, int *max_score){
	uint32_t W = a->W, j, c, i, nchunk = (W + 31) / 32, best_lane, best_chunk, x, y, pos;
	int32_t Max[BSO_LANES], Scr[BSO_LANES];
	uint32_t Idx[BSO_LANES];
	int umax, uscr;
	for(j=0;j<BSO_LANES;j++){ Max[j] = BSO_SCORE_MIN; Scr[j] = r.ub[j]; Idx[j] = j; }
	for(c=0;c<nchunk;c++){
		uint32_t lo = c * 32, hi = lo + 32 < W ? lo + 32 : W;
		for(j=0;j<BSO_LANES;j++){
			int run = 0, mx = -32767, h; /* int16 lanes in the reference; |run| <= 32*128 never saturates */
			for(i=lo;i<hi;i++){ run += r.u[j * W + i]; if(run > mx) mx = run; }
			h = Scr[j] + mx;
			if(h > Max[j]){ Max[j] = h; Idx[j] = j | (c << 8); }
			Scr[j] += run;
		}
	}
	/* lanes j, j+4, j+8, j+12 are folded pairwise keeping the lower lane on ties (:3264-3266) */
	for(j=0;j<4;j++){
		int32_t m0 = Max[j], m1 = Max[j + 8];
		uint32_t i0 = Idx[j], i1 = Idx[j + 8];
		if(Max[j + 4] > m0){ m0 = Max[j + 4]; i0 = Idx[j + 4]; }
		if(Max[j + 12] > m1){ m1 = Max[j + 12]; i1 = Idx[j + 12]; }
		if(m1 > m0){ m0 = m1; i0 = i1; }
		Max[j] = m0; Idx[j] = i0;
	}
	*max_score = Max[0]; x = 0;
	for(j=1;j<4;j++) if(Max[j] > *max_score){ *max_score = Max[j]; x = j; }
	best_lane = Idx[x] & 0xFF; best_chunk = Idx[x] >> 8;
	x = best_chunk * 32; y = (best_chunk + 1) * 32 < W ? (best_chunk + 1) * 32 : W;
	pos = x; umax = BSO_SCORE_MIN; uscr = 0;
	for(;x<y;x++){
		uscr += r.u[best_lane * W + x];
		if(uscr > umax){ pos = x; umax = uscr; }
	}
	return best_lane * W + pos;
}

/* bsalign.h:3331-3349 */
static int bso_band_mov(bso_epi8_t *a, const int32_t *ub, uint32_t tidx, uint32_t qoff){
	uint32_t W = a->W, i;
	int noisy = 0;
	uint32_t nz;
	if(tidx <= W * BSO_LANES / 4) return 0;
	if(qoff + W * BSO_LANES >= a->qlen) return 0;
	for(i=1;i<=BSO_LANES;i++) noisy += ub[i] < ub[i - 1] ? ub[i - 1] - ub[i] : ub[i] - ub[i - 1];
	nz = ((uint32_t)(noisy / BSO_LANES)) / W * BSO_LANES / 2; /* int / u4i promotes to unsigned in the reference */
	noisy = (int)(16u > nz ? 16u : nz);
	if(ub[0] + noisy < ub[BSO_LANES]) return 2;
	if(ub[0] > ub[BSO_LANES] + noisy) return 0;
	return 1;
}

typedef struct { uint32_t *buf; uint32_t cap, n; uint32_t run; int on; } bso_cig_t;

/* bsalign.h:409-417 */
static void bso_cig_push(bso_epi8_t *a, bso_cig_t *cg, uint32_t op, uint32_t sz){
	if(op == (cg->run & 0xf)){ cg->run += sz << 4; return; }
	if(cg->run){
		if(cg->on){ if(cg->n < cg->cap) cg->buf[cg->n] = cg->run; else if(a) a->err |= BSO_ERR_CIGCAP; }
		cg->n++;
	}
	cg->run = sz << 4 | op;
}
static void bso_cig_flush(bso_epi8_t *a, bso_cig_t *cg){
	if(cg->run){
		if(cg->on){ if(cg->n < cg->cap) cg->buf[cg->n] = cg->run; else if(a) a->err |= BSO_ERR_CIGCAP; }
		cg->n++;
	}
	cg->run = 0;
}

/* H(col, row) from the trace (bsalign.h:3199-3202) */
static int bso_mtx_score(bso_epi8_t *a, int64_t row, int64_t col){
	if(row < -1 || row >= (int64_t)a->tlen){ a->err |= BSO_ERR_RANGE; return BSO_SCORE_MIN; }
	return bso_getscore(a, bso_trace_row(a, row), col - a->begs[row + 1]);
}

/* bsalign.h:3667-3702 and :3704-3852: traceback by re-deriving each step from score differences */
static void bso_backcal(bso_epi8_t *a, int32_t *rs, bso_cig_t *cg){
	int qb = rs[2], tb = rs[4]; /* inclusive end cell */
	int mat = 0, mis = 0, ins = 0, del = 0, aln = 0;
	int Hcur, Hprev = 0, pend = 0, prior = 0, bw = (int)a->bw, pw = a->pw;
	int64_t guard = 0, guard_max = 8 * ((int64_t)a->qlen + a->tlen) + 64;
	rs[2] = qb + 1; rs[4] = tb + 1;
	Hcur = bso_mtx_score(a, tb, qb);
	while(1){
		if(++guard > guard_max){ a->err |= BSO_ERR_LOOP; break; }
		if((pend & 0xf) == 2 || (pend & 0xf) == 4){ /* inside a deletion run (piece 1: op 2, piece 2: op 4) */
			int len = pend >> 4, cost;
			Hprev = bso_mtx_score(a, tb, qb);
			cost = ((pend & 0xf) == 2) ? a->go1 + len * a->ge1 : a->go2 + len * a->ge2;
			if(Hprev + cost == Hcur){
				bso_cig_push(a, cg, 2, len);
				del += len; aln += len;
				Hcur = Hprev; pend = 0;
			} else {
				pend += 1 << 4; tb--;
				continue;
			}
		}
		if(qb < 0 || tb < 0) break;
		if(qb == a->begs[tb]){ /* begs[] is shifted by one: this is the band start of row tb-1 (:3761) */
			if(qb){
				Hprev = a->UB[(int64_t)tb * 17]; /* ub[0] of row tb-1 */
				prior = 0;
			} else if(a->mode == BSO_MODE_OVERLAP || tb == 0) Hprev = 0;
			else if(pw < 2) Hprev = a->go1 + a->ge1 * tb;
			else { int c1 = a->go1 + a->ge1 * tb, c2 = a->go2 + a->ge2 * tb; Hprev = c1 > c2 ? c1 : c2; }
		} else if(qb - a->begs[tb] <= bw){
			Hprev = bso_mtx_score(a, tb - 1, qb - 1);
		} /* else: right of row tb-1's band; the reference reads past the row but never uses the value (:3671) */
		{
			int x = qb - a->begs[tb], bt, s, h;
			int8_t u = 0, e = 0, q = 0;
			if(x >= 0 && x < bw){
				bso_row_t r = bso_trace_row(a, tb - 1);
				u = r.u[x]; e = r.e ? r.e[x] : (int8_t)(a->go1 + a->ge1); q = r.q ? r.q[x] : 0;
			}
			s = a->mtx[a->qseq[qb] * 4 + a->tseq[tb]];
			h = Hcur - Hprev;
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
				if(a->qseq[qb] == a->tseq[tb]) mat++; else mis++;
				qb--; tb--; aln++;
				bso_cig_push(a, cg, 0, 1);
				Hcur = Hprev;
			} else if(bt == 1){
				if(qb <= 0){
					bso_cig_push(a, cg, 1, 1);
					Hcur = Hprev;
					qb--; ins++; aln++;
				} else {
					int sz, t, Hl;
					for(sz=1;sz+a->begs[tb + 1]<=qb;sz++){
						t = a->go1 + sz * a->ge1;
						if(pw == 2){ int t2 = a->go2 + sz * a->ge2; if(t2 > t) t = t2; }
						Hl = bso_mtx_score(a, tb, qb - sz);
						if(Hl + t == Hcur){
							bso_cig_push(a, cg, 1, sz);
							Hcur = Hl; qb -= sz; ins += sz; aln += sz;
							break;
						}
					}
					/* no match: the reference re-enters the same state forever; the guard above ends it */
				}
			} else {
				pend = (1 << 4) | bt;
				tb--;
				continue;
			}
		}
	}
	if(a->mode == BSO_MODE_OVERLAP){
		bso_cig_flush(a, cg);
	} else {
		uint32_t op = 0, sz = 0;
		if(qb >= 0){ op = 1; sz = qb + 1; ins += sz; qb = -1; }
		else if(tb >= 0){ op = 2; sz = tb + 1; del += sz; tb = -1; }
		aln += sz;
		bso_cig_push(a, cg, op, sz);
		bso_cig_flush(a, cg);
	}
	rs[1] = qb + 1; rs[3] = tb + 1;
	rs[5] = mat; rs[6] = mis; rs[7] = ins; rs[8] = del; rs[9] = aln;
	if(cg->on){ /* :3850 */
		uint32_t i, n = cg->n < cg->cap ? cg->n : cg->cap;
		for(i=0;i<n/2;i++){ uint32_t t = cg->buf[i]; cg->buf[i] = cg->buf[n - 1 - i]; cg->buf[n - 1 - i] = t; }
	}
}

/*
 * bsalign.h:3854-4050.  result = {score,qb,qe,tb,te,mat,mis,ins,del,aln}.  Returns error flags (0 = ok).
 * dump_* (optional) receive per-row band offsets, anchors and linear u/e/q rows for debugging.
 */
int bso_epi8_pairwise_ex(const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen, int mode, uint32_t bandwidth,
		const int8_t mtx[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		int32_t *rs, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar,
		int32_t *dump_begs, int32_t *dump_ub, int8_t *dump_u, int8_t *dump_e, int8_t *dump_q){
	bso_epi8_t A, *a = &A;
	bso_row_t prev, tmp, cur;
	bso_cig_t cg;
	uint32_t i, bw, rbeg, mov;
	int k, rh, rbx, rby, rbz, score, max_score;
	int8_t *scratch;
	int32_t tub[17];
	memset(a, 0, sizeof(A));
	memset(rs, 0, 10 * sizeof(int32_t));
	if(ncigar) *ncigar = 0;
	if(qlen == 0 || tlen == 0) return 0; /* the reference has no guard here; callers never pass empties */
	bw = bandwidth ? bandwidth : qlen;
	bw = (bw + BSO_LANES - 1) / BSO_LANES * BSO_LANES;
	a->qlen = qlen; a->tlen = tlen; a->bw = bw; a->W = bw / BSO_LANES; a->mode = mode & 3;
	a->qseq = qseq; a->tseq = tseq; a->mtx = mtx; a->go1 = go1; a->ge1 = ge1; a->go2 = go2; a->ge2 = ge2;
	a->pw = bso_epi8_piecewise(go1, ge1, go2, ge2, bw);
	a->smax = -127; a->smin = 127;
	for(k=0;k<16;k++){ if(mtx[k] > a->smax) a->smax = mtx[k]; if(mtx[k] < a->smin) a->smin = mtx[k]; }
	a->U = malloc((size_t)bw * (tlen + 1));
	a->E = a->pw ? malloc((size_t)bw * (tlen + 1)) : NULL;
	a->Q = a->pw == 2 ? malloc((size_t)bw * (tlen + 1)) : NULL;
	a->UB = malloc(sizeof(int32_t) * 17 * (size_t)(tlen + 1));
	a->begs = malloc(sizeof(int32_t) * (size_t)(tlen + 1));
	scratch = malloc((size_t)bw * 3);
	tmp.u = scratch; tmp.e = scratch + bw; tmp.q = scratch + 2 * (size_t)bw; tmp.ub = tub;
	prev = bso_trace_row(a, -1);
	bso_row_init(a, prev, a->smax, a->smin);
	a->begs[0] = 0;
	rs[0] = BSO_SCORE_MIN;
	rbeg = 0; mov = 0;
	for(i=0;i<tlen;i++){
		uint8_t tb = tseq[i];
		if(mov && rbeg + bw < qlen){
			int lim = (int)qlen - (int)(rbeg + bw);
			if(lim < 0) lim = 0;
			if((uint32_t)lim < mov) mov = (uint32_t)lim;
			rbeg += mov;
			rh = bso_getscore(a, prev, (int64_t)mov - 1);
		} else {
			mov = 0;
			if(rbeg) rh = BSO_SCORE_MIN;
			else if(a->mode == BSO_MODE_OVERLAP || i == 0) rh = 0;
			else if(a->pw < 2) rh = (int)((uint32_t)go1 + (uint32_t)ge1 * i);
			else {
				uint32_t c1 = (uint32_t)go1 + (uint32_t)ge1 * i, c2 = (uint32_t)go2 + (uint32_t)ge2 * i; /* unsigned compare, :3944 */
				rh = (int)(c1 > c2 ? c1 : c2);
			}
		}
		bso_row_shift(a, prev, tmp, mov, a->smax, a->smin);
		cur = bso_trace_row(a, i);
		bso_row_cal(a, rbeg, tb, tmp, cur, rh);
		rbx = bso_band_mov(a, cur.ub, i, rbeg);
		if(a->mode == BSO_MODE_GLOBAL){
			int tq = (int)(tlen / qlen);
			rbz = 2 * (tq > 1 ? tq : 1);
			rby = (int)((1.0 * i / tlen) * qlen);
			if((int64_t)rbeg + rbz * (int64_t)(tlen - i - 1) + (int64_t)bw <= (int64_t)(uint32_t)(qlen + (uint32_t)rbz - 1)){
				uint32_t rem = tlen - i - 1;
				mov = 1 + ((qlen - (rbeg + bw)) / (rem > 1 ? rem : 1));
			} else if((int)rbeg < rby - (int)bw){
				mov = rbx + 1;
			} else if((int)rbeg > rby){
				mov = rbx - 1 > 0 ? rbx - 1 : 0;
			} else mov = rbx;
		} else mov = rbx;
		a->begs[i + 1] = rbeg;
		if(a->mode != BSO_MODE_GLOBAL && rbeg + bw >= qlen){
			score = bso_getscore(a, cur, (int64_t)qlen - 1 - rbeg);
			if(score > rs[0]){ rs[0] = score; rs[2] = qlen - 1; rs[4] = i; }
		}
		prev = cur;
	}
	if(a->mode == BSO_MODE_GLOBAL){
		rs[0] = bso_getscore(a, prev, (int64_t)qlen - 1 - rbeg);
		rs[2] = qlen - 1; rs[4] = tlen - 1;
	} else {
		uint32_t rmax = bso_row_max(a, prev, &max_score);
		if(max_score > rs[0]){ rs[0] = max_score; rs[2] = rbeg + rmax; rs[4] = tlen - 1; }
	}
	if(dump_begs){
		for(i=0;i<tlen;i++){
			dump_begs[i] = a->begs[i + 1];
			memcpy(dump_ub + (size_t)i * 17, a->UB + (size_t)(i + 1) * 17, 17 * sizeof(int32_t));
			memcpy(dump_u + (size_t)i * bw, a->U + (size_t)(i + 1) * bw, bw);
			if(dump_e){ if(a->E) memcpy(dump_e + (size_t)i * bw, a->E + (size_t)(i + 1) * bw, bw); else memset(dump_e + (size_t)i * bw, 0, bw); }
			if(dump_q){ if(a->Q) memcpy(dump_q + (size_t)i * bw, a->Q + (size_t)(i + 1) * bw, bw); else memset(dump_q + (size_t)i * bw, 0, bw); }
		}
	}
	cg.buf = cigar; cg.cap = cigar_cap; cg.n = 0; cg.run = 0; cg.on = cigar != NULL;
	bso_backcal(a, rs, &cg);
	if(ncigar) *ncigar = cg.n;
	free(a->U); free(a->E); free(a->Q); free(a->UB); free(a->begs); free(scratch);
	return a->err;
}

int bso_epi8_pairwise(const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen, int mode, uint32_t bandwidth,
		const int8_t mtx[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		int32_t *rs, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar){
	return bso_epi8_pairwise_ex(qseq, qlen, tseq, tlen, mode, bandwidth, mtx, go1, ge1, go2, ge2, rs, cigar, cigar_cap, ncigar, NULL, NULL, NULL, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * Batch runners (pthread pool over independent pairs) -- same argument layout as oracle/ref_harness.c
 * ------------------------------------------------------------------------------------------------ */
int bso_edit_pairwise(const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen, int mode, uint32_t bandwidth,
		int32_t *rs, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar);
int bso_kmer_edit_pairwise(uint32_t ksz, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen,
		int32_t *rs, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar);

typedef struct {
	int kind;
	uint64_t n;
	const uint8_t *seqs;
	const uint64_t *qoff, *toff;
	const uint32_t *qlen, *tlen;
	int mode;
	uint32_t bandwidth;
	const int8_t *matrix;
	int8_t go1, ge1, go2, ge2;
	int32_t *results;
	uint32_t *cigars;
	const uint64_t *cgoff;
	uint32_t *ncigar;
	volatile uint64_t next;
	int repeat;
	volatile int err;
	int32_t *errs;
} bso_job_t;

static void* bso_worker(void *arg){
	bso_job_t *job = (bso_job_t*)arg;
	uint64_t i, j;
	int r, err = 0;
	while(1){
		i = __sync_fetch_and_add(&job->next, 16);
		if(i >= job->n) break;
		for(j=i;j<i+16&&j<job->n;j++){
			uint32_t *cg = (job->cigars && job->cgoff) ? job->cigars + job->cgoff[j] : NULL;
			uint32_t cap = (job->cigars && job->cgoff) ? (uint32_t)(job->cgoff[j + 1] - job->cgoff[j]) : 0;
			uint32_t ncg = 0;
			int perr = 0;
			for(r=0;r<job->repeat;r++){
				if(job->kind == 0){
					perr |= bso_epi8_pairwise(job->seqs + job->qoff[j], job->qlen[j], job->seqs + job->toff[j], job->tlen[j], job->mode, job->bandwidth,
						job->matrix, job->go1, job->ge1, job->go2, job->ge2, job->results + j * 10, cg, cap, &ncg);
				} else if(job->kind == 1){
					perr |= bso_edit_pairwise(job->seqs + job->qoff[j], job->qlen[j], job->seqs + job->toff[j], job->tlen[j], job->mode, job->bandwidth,
						job->results + j * 10, cg, cap, &ncg);
				} else {
					perr |= bso_kmer_edit_pairwise(job->bandwidth, job->seqs + job->qoff[j], job->qlen[j], job->seqs + job->toff[j], job->tlen[j],
						job->results + j * 10, cg, cap, &ncg);
				}
			}
			err |= perr;
			if(job->errs) job->errs[j] = perr;
			if(job->ncigar) job->ncigar[j] = ncg;
		}
	}
	if(err) __sync_fetch_and_or(&job->err, err);
	return NULL;
}

static int bso_run(bso_job_t *job, int nthreads){
	pthread_t *tids;
	int i;
	if(nthreads < 1) nthreads = 1;
	if(job->repeat < 1) job->repeat = 1;
	job->next = 0; job->err = 0;
	if(nthreads == 1){ bso_worker(job); return job->err; }
	tids = malloc(sizeof(pthread_t) * nthreads);
	for(i=0;i<nthreads;i++) pthread_create(tids + i, NULL, bso_worker, job);
	for(i=0;i<nthreads;i++) pthread_join(tids[i], NULL);
	free(tids);
	return job->err;
}

int bso_epi8_batch_ex(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat, int32_t *errs){
	bso_job_t job;
	memset(&job, 0, sizeof(job));
	job.kind = 0; job.n = n; job.seqs = seqs; job.qoff = qoff; job.qlen = qlen; job.toff = toff; job.tlen = tlen;
	job.mode = mode; job.bandwidth = bandwidth; job.matrix = matrix; job.go1 = go1; job.ge1 = ge1; job.go2 = go2; job.ge2 = ge2;
	job.results = results; job.cigars = cigars; job.cgoff = cgoff; job.ncigar = ncigar; job.repeat = repeat; job.errs = errs;
	return bso_run(&job, nthreads);
}

int bso_epi8_batch(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat){
	return bso_epi8_batch_ex(n, seqs, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, results, cigars, cgoff, ncigar, nthreads, repeat, NULL);
}

int bso_edit_batch_ex(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat, int32_t *errs);

int bso_edit_batch(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat){
	return bso_edit_batch_ex(n, seqs, qoff, qlen, toff, tlen, mode, bandwidth, results, cigars, cgoff, ncigar, nthreads, repeat, NULL);
}

int bso_edit_batch_ex(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat, int32_t *errs){
	bso_job_t job;
	memset(&job, 0, sizeof(job));
	job.kind = 1; job.errs = errs; job.n = n; job.seqs = seqs; job.qoff = qoff; job.qlen = qlen; job.toff = toff; job.tlen = tlen;
	job.mode = mode; job.bandwidth = bandwidth;
	job.results = results; job.cigars = cigars; job.cgoff = cgoff; job.ncigar = ncigar; job.repeat = repeat;
	return bso_run(&job, nthreads);
}

/* k-mer guided edit (kind 2; the k-mer size travels in the bandwidth field); cigar capacity per pair: 2 * (qlen + tlen) + 4 words */
int bso_kmer_edit_batch(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		uint32_t ksz, int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat, int32_t *errs){
	bso_job_t job;
	memset(&job, 0, sizeof(job));
	job.kind = 2; job.errs = errs; job.n = n; job.seqs = seqs; job.qoff = qoff; job.qlen = qlen; job.toff = toff; job.tlen = tlen;
	job.mode = 0; job.bandwidth = ksz;
	job.results = results; job.cigars = cigars; job.cgoff = cgoff; job.ncigar = ncigar; job.repeat = repeat;
	return bso_run(&job, nthreads);
}

/* ------------------------------------------------------------------------------------------------
 * Edit-distance kernel (bsalign.h:612-1206), restated per cell.
 *
 * The reference keeps u(x,y) = H(x,y) - H(x-1,y) in two bit-planes (minus, plus) of 64-bit words,
 * striped over 64 lanes, and resolves the left-to-right dependency with a re-pass loop that stops at the
 * exact fix-point (bsalign.h:784-809).  That fix-point is the plain sequential recurrence
 *     h  = (q[x] != t[y]) && u(x,y-1) != -1 && v(x-1,y) != -1      (h = H(x,y) - H(x-1,y-1), 0 or 1)
 *     u' = h - v(x-1,y),   v' = h - u(x,y-1)
 * with v = +1 entering the band's first cell (0 in OVERLAP mode), which is what is evaluated here.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
	uint32_t qlen, tlen, bw, W;
	int mode;
	const uint8_t *qseq, *tseq;
	int8_t *U;       /* (tlen+1) rows of bw cells; row 0 is the init row (all +1), row y+1 is target row y */
	uint32_t *begs;  /* tlen+1; begs[y+1] = band start of row y */
	int err;
} bso_edit_t;

static inline int bso_edit_u(bso_edit_t *a, int64_t trow, int64_t pos){
	if(pos < 0 || pos >= (int64_t)a->bw || trow < 0 || trow > (int64_t)a->tlen){ a->err |= BSO_ERR_RANGE; return 0; }
	return a->U[trow * (int64_t)a->bw + pos];
}

/* bsalign.h:813-963: arg-min of the last row with the SSE code's block/chunk tie-break order (EXTEND only) */
static int bso_edit_rowmin(bso_edit_t *a, int sbeg0, const int8_t *u, uint32_t *whence){
	uint32_t W = a->W, blk, l, ib, ie, i, pmin = 0;
	int sbeg = sbeg0, smin = sbeg0;
	for(blk=0;blk<4;blk++){
		int tot[16], mm[16], cand[16], sc;
		uint32_t pp[16], st;
		for(l=0;l<16;l++){
			const int8_t *lane = u + (size_t)(blk * 16 + l) * W;
			int hh = 0;
			mm[l] = 0; pp[l] = 0;
			for(ib=0;ib<W;ib=ie){
				int h = 0, m = 0, d;
				uint32_t pz = 0;
				ie = ib + 124 < W ? ib + 124 : W;
				for(i=ib;i<ie;i++){
					h += lane[i];
					if(m > h){ m = h; pz = i - ib; }
				}
				d = hh + m;
				if(mm[l] > d){ mm[l] = d; pp[l] = pz + ib; }
				hh += h;
			}
			tot[l] = hh;
		}
		for(l=0;l<16;l++){ cand[l] = sbeg + mm[l]; sbeg += tot[l]; }
		sc = cand[0]; st = 0;
		for(l=1;l<16;l++) if(sc > cand[l]){ sc = cand[l]; st = l; }
		if(sc >= smin) continue;
		smin = sc;
		pmin = (blk * 16 + st) * W + pp[st];
	}
	if(whence) *whence = pmin;
	return smin;
}

/* bsalign.h:1046-1206 (driver), :653-721 (row init / shift), :766-810 (row), :965-1044 (backtrace) */
int bso_edit_pairwise(const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen, int mode, uint32_t bandwidth,
		int32_t *rs, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar){
	bso_edit_t A, *a = &A;
	bso_cig_t cg;
	uint32_t i, p, bw, q64, rbeg = 0, prev_beg = 0;
	int type = mode & 3, sbeg = 0, smin = 0x7FFFFFFF, rx, ry, srow;
	int8_t *prev, *cur, *tmp;
	memset(rs, 0, 10 * sizeof(int32_t));
	if(ncigar) *ncigar = 0;
	if(qlen == 0 || tlen == 0) return 0; /* :1051-1054 */
	q64 = (qlen + 63) / 64 * 64;
	if(type == BSO_MODE_OVERLAP || type == BSO_MODE_EXTEND) bw = q64; /* :1055-1067 */
	else {
		bw = (bandwidth + 63) / 64 * 64;
		if(bw == 0 || bw > qlen) bw = q64;
		if(bw < qlen && bw < ((qlen + tlen - 1) / tlen) + 1) bw = ((qlen + tlen - 1) / tlen + 1 + 63) / 64 * 64;
	}
	if(bw > q64) return BSO_ERR_RANGE; /* qlen % 64 == 0 with a tiny target: the reference's band start underflows (:1114) */
	memset(a, 0, sizeof(A));
	a->qlen = qlen; a->tlen = tlen; a->bw = bw; a->W = bw / 64; a->mode = type; a->qseq = qseq; a->tseq = tseq;
	a->U = malloc((size_t)bw * (tlen + 1));
	a->begs = malloc(sizeof(uint32_t) * (size_t)(tlen + 1));
	tmp = malloc(bw);
	memset(a->U, 1, bw); /* init row: u = +1 everywhere (:653-656) */
	a->begs[0] = 0;
	rx = qlen - 1; ry = tlen - 1;
	for(i=0;i<tlen;i++){
		uint32_t movx;
		int v;
		prev = a->U + (size_t)i * bw;
		cur = a->U + (size_t)(i + 1) * bw;
		if(type == BSO_MODE_OVERLAP || type == BSO_MODE_EXTEND) rbeg = 0;
		else { /* fixed diagonal band (:1112-1114) */
			rbeg = (uint32_t)(((uint64_t)i * qlen) / tlen);
			rbeg = (rbeg < bw / 2) ? 0 : rbeg - bw / 2;
			if(rbeg + bw > q64) rbeg = q64 - bw;
		}
		a->begs[i + 1] = rbeg;
		movx = rbeg - prev_beg;
		/* shift (:658-721): sbeg follows H(rbeg-1, y) */
		if(type == BSO_MODE_OVERLAP){
			sbeg = 0;
			memcpy(tmp, prev, bw);
		} else {
			uint32_t mv = movx < bw ? movx : bw;
			/* bsalign.h:704-713 offsets its destination pointers twice when the shift spans whole 64-lane
			 * cycles plus a remainder; the reference then computes on stale scratch words.  Out of domain. */
			if(movx < bw && movx >= a->W && movx % a->W) a->err |= BSO_ERR_REFBUG;
			for(p=0;p<mv;p++) sbeg += prev[p];
			sbeg++;
			for(p=0;p<bw;p++) tmp[p] = (p + movx < bw) ? prev[p + movx] : 1;
		}
		/* the row */
		v = (type == BSO_MODE_OVERLAP) ? 0 : 1;
		for(p=0;p<bw;p++){
			uint32_t x = rbeg + p;
			int match = (x < qlen) && (qseq[x] == tseq[i]);
			int u = tmp[p];
			int h = (!match && u != -1 && v != -1) ? 1 : 0;
			cur[p] = (int8_t)(h - v);
			v = h - u;
		}
		if(type == BSO_MODE_OVERLAP || type == BSO_MODE_EXTEND){ /* :1124-1139: H(qlen-1, i) */
			srow = sbeg;
			for(p=0;p<bw&&rbeg+p<qlen;p++) srow += cur[p];
			if(srow < smin){ smin = srow; rx = qlen - 1; ry = i; }
		}
		prev_beg = rbeg;
	}
	cur = a->U + (size_t)tlen * bw;
	if(type == BSO_MODE_EXTEND){
		uint32_t k = 0;
		srow = bso_edit_rowmin(a, sbeg, cur, &k);
		if(srow < smin){ smin = srow; rx = k; ry = tlen - 1; }
	}
	/* backtrace (:965-1044) */
	{
		int x = rx, y = ry, mat = 0, mis = 0, ins = 0, del = 0;
		uint32_t op;
		int64_t guard = 0;
		cg.buf = cigar; cg.cap = cigar_cap; cg.n = 0; cg.run = 0; cg.on = cigar != NULL;
		rs[2] = x + 1; rs[4] = y + 1;
		while(x >= 0 && y >= 0){
			if(++guard > 4 * ((int64_t)qlen + tlen) + 64){ a->err |= BSO_ERR_LOOP; break; }
			if(qseq[x] == tseq[y]){ mat++; op = 0; x--; y--; }
			else if(bso_edit_u(a, y + 1, (int64_t)x - a->begs[y + 1]) == 1){ ins++; op = 1; x--; }
			else if(bso_edit_u(a, y, (int64_t)x - a->begs[y]) == -1){ del++; op = 2; y--; }
			else { mis++; op = 0; x--; y--; }
			/* :1009-1014 is the same run-length merge as bso_cig_push with length 1 */
			if(op == (cg.run & 0xf)) cg.run += 0x10;
			else { bso_cig_flush(NULL, &cg); cg.run = 0x10 | op; }
		}
		rs[1] = x + 1; rs[3] = y + 1;
		if(rs[1]){
			op = 1;
			if(op == (cg.run & 0xf)) cg.run += 0x10 * (uint32_t)rs[1];
			else { bso_cig_flush(NULL, &cg); cg.run = (0x10 * (uint32_t)rs[1]) | op; }
			ins += rs[1]; rs[1] = 0;
		}
		if((type == BSO_MODE_GLOBAL || type == BSO_MODE_EXTEND) && rs[3]){
			op = 2;
			if(op == (cg.run & 0xf)) cg.run += 0x10 * (uint32_t)rs[3];
			else { bso_cig_flush(NULL, &cg); cg.run = (0x10 * (uint32_t)rs[3]) | op; }
			del += rs[3]; rs[3] = 0;
		}
		rs[5] = mat; rs[6] = mis; rs[7] = ins; rs[8] = del; rs[9] = mat + mis + ins + del;
		bso_cig_flush(NULL, &cg);
		if(cg.on){
			uint32_t n = cg.n < cg.cap ? cg.n : cg.cap, k;
			if(cg.n > cg.cap) a->err |= BSO_ERR_CIGCAP;
			for(k=0;k<n/2;k++){ uint32_t t = cg.buf[k]; cg.buf[k] = cg.buf[n - 1 - k]; cg.buf[n - 1 - k] = t; }
		}
		if(ncigar) *ncigar = cg.n;
	}
	/* score (:1189-1203) */
	if(type == BSO_MODE_OVERLAP) rs[0] = smin + rs[4] - rs[3];
	else if(type == BSO_MODE_EXTEND) rs[0] = smin;
	else {
		int sc = sbeg;
		for(p=0;p<bw&&rbeg+p<qlen;p++) sc += cur[p];
		rs[0] = sc;
	}
	free(a->U); free(a->begs); free(tmp);
	return a->err;
}


/* ------------------------------------------------------------------------------------------------
 * POA read-vs-graph sweep: align_rd_bspoacore (bspoa.h:2515-2618) with dpalign_row_update_bspoa
 * (bspoa.h:2232-2261), dpalign_row_merge_bspoa (bspoa.h:2263-2272) and
 * banded_striped_epi8_seqalign_piecex_row_merge (bsalign.h:2474-2616).
 *
 * The sub-graph arrives as CSR over LOCAL node ids (0..nnode-1, the order of g->sels), out-edges in the
 * reference's own list order and already filtered to selected nodes (bspoa.h:2538).  Rows are kept in
 * linear band order here; par = {bandwidth, alnmode, M, X, O, E, Q, P, T, refbonus}.
 * ------------------------------------------------------------------------------------------------ */
static inline int16_t sat16(int v){ return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }

/* bsalign.h:2474-2616: out = cell-wise max of rows a and b in absolute space, re-differenced.  out may alias b. */
static void bso_row_merge(uint32_t W, int pw, bso_row_t a, bso_row_t b, bso_row_t out){
	uint32_t j, i, ib, ie;
	int k;
	for(j=0;j<BSO_LANES;j++){
		int32_t sa = a.ub[j], sb = b.ub[j];
		for(ib=0;ib<W;ib=ie){
			int d, xa, xb;
			int16_t ta, tb, mp, mc;
			ie = ib + 256 < W ? ib + 256 : W;
			d = sa - sb;
			if(d < -0x7FFF) d = -0x7FFF;
			if(d > 0x7FFF) d = 0x7FFF;
			xa = d >> 1; /* arithmetic shift (:2509) */
			xb = xa - d;
			sa -= xa; sb -= xb;
			ta = sat16(xa); tb = sat16(xb);
			mp = ta > tb ? ta : tb;
			for(i=ib;i<ie;i++){
				uint32_t p = j * W + i;
				int16_t ya, yb, mx;
				ta = sat16(ta + a.u[p]); tb = sat16(tb + b.u[p]);
				mc = ta > tb ? ta : tb;
				if(pw >= 1){ ya = sat16(ta + a.e[p]); yb = sat16(tb + b.e[p]); mx = ya > yb ? ya : yb; out.e[p] = sat8(sat16(mx - mc)); }
				if(pw == 2){ ya = sat16(ta + a.q[p]); yb = sat16(tb + b.q[p]); mx = ya > yb ? ya : yb; out.q[p] = sat8(sat16(mx - mc)); }
				out.u[p] = sat8(sat16(mc - mp));
				mp = mc;
			}
			sa += ta; sb += tb;
		}
	}
	for(k=0;k<=BSO_LANES;k++) out.ub[k] = a.ub[k] > b.ub[k] ? a.ub[k] : b.ub[k];
}

int bso_poa_sweep(const int32_t *par, const uint8_t *query, uint32_t slen, uint32_t nnode,
		const uint8_t *base, const uint8_t *bonus, const int32_t *rpos, const int32_t *nct,
		const int32_t *eoff, const int32_t *edst, uint32_t head, uint32_t tail,
		int8_t *rows, int32_t *ubs, uint8_t *done, int32_t *best, uint64_t *ops){
	bso_epi8_t A, *a = &A;
	uint32_t bw = (uint32_t)par[0], W = bw / BSO_LANES, sp = 0, n, k;
	int alnmode = par[1] & 3, M = par[2], X = par[3], O = par[4], E = par[5], Q = par[6], P = par[7], T = par[8], refbonus = par[9];
	int8_t mtx[2][16], *scratch;
	int32_t sub[2][17], *mpos;
	uint32_t *stack, *vst;
	bso_row_t blk0, blk1;
	int maxscr = BSO_SCORE_MIN, maxidx = -1, maxoff = -1;
	uint64_t nupd = 0, nmrg = 0;
	if(bw == 0 || bw % BSO_LANES || head >= nnode || tail >= nnode) return BSO_ERR_RANGE;
	memset(a, 0, sizeof(A));
	a->qlen = slen; a->bw = bw; a->W = W; a->mode = alnmode; a->qseq = query;
	a->go1 = (int8_t)O; a->ge1 = (int8_t)E; a->go2 = (int8_t)Q; a->ge2 = (int8_t)P;
	a->pw = bso_epi8_piecewise(a->go1, a->ge1, a->go2, a->ge2, bw);
	a->smax = (int8_t)(M + refbonus + 1); a->smin = (int8_t)X; /* bspoa.h:2226, 2240 */
	for(k=0;k<16;k++){ /* bsalign.h:323 */
		int same = ((k ^ (k >> 2)) & 3) == 0;
		mtx[0][k] = (int8_t)(same ? M : X);
		mtx[1][k] = (int8_t)(same ? M + refbonus : X);
	}
	scratch = malloc((size_t)bw * 6);
	blk0.u = scratch; blk0.e = scratch + bw; blk0.q = scratch + 2 * (size_t)bw; blk0.ub = sub[0];
	blk1.u = scratch + 3 * (size_t)bw; blk1.e = blk1.u + bw; blk1.q = blk1.u + 2 * (size_t)bw; blk1.ub = sub[1];
	mpos = malloc(sizeof(int32_t) * nnode); stack = malloc(sizeof(uint32_t) * (nnode + 1)); vst = calloc(nnode, sizeof(uint32_t));
	for(n=0;n<nnode;n++) mpos[n] = 0x7FFFFFFF - 1; /* MAX_B4 - 1, bspoa.h:2525 */
	memset(done, 0, nnode);
	#define NODE_ROW(r, n_) do { (r).u = rows + (size_t)(n_) * 3 * bw; (r).e = (r).u + bw; (r).q = (r).u + 2 * (size_t)bw; (r).ub = ubs + (size_t)(n_) * 17; } while(0)
	{ /* bspoa.h:2224-2226: the head row */
		bso_row_t r; NODE_ROW(r, head);
		memset(r.u, 0, 3 * (size_t)bw);
		bso_row_init(a, r, a->smax, a->smin);
		if(a->pw == 0){ /* row_init leaves e untouched without a gap-open cost */ }
		done[head] = 1;
	}
	mpos[head] = -1;
	stack[sp++] = head;
	while(sp){
		uint32_t u = stack[--sp];
		int32_t ei;
		bso_row_t ur; NODE_ROW(ur, u);
		for(ei=eoff[u];ei<eoff[u+1];ei++){
			uint32_t v = (uint32_t)edst[ei];
			if(mpos[u] + 1 < mpos[v]) mpos[v] = mpos[u] + 1;
			if(v == tail){ /* bspoa.h:2548-2580 */
				int smax, moff = (int)((int)slen < rpos[u] + (int)bw ? (int)slen : rpos[u] + (int)bw) - 1;
				smax = bso_getscore(a, ur, (int64_t)moff - rpos[u]);
				if((int)slen > moff + 1){
					int rem = (int)slen - moff - 1;
					if(a->pw < 2) smax += O + E * rem;
					else { int c1 = O + E * rem, c2 = Q + P * rem; smax += c1 > c2 ? c1 : c2; }
				}
				smax += T;
				if(smax > maxscr){ maxscr = smax; maxidx = (int)u; maxoff = moff; }
				if(alnmode == BSO_MODE_OVERLAP){
					uint32_t rmax = bso_row_max(a, ur, &smax);
					if(smax > maxscr){ maxscr = smax; maxidx = (int)u; maxoff = (int)rmax + rpos[u]; }
				}
				vst[v]++;
			} else {
				bso_row_t vr, dst;
				int rh, kprof = (base[v] == base[u]) * 2 + bonus[v];
				uint32_t q1 = (uint32_t)rpos[u], q2 = (uint32_t)rpos[v];
				NODE_ROW(vr, v);
				dst = vst[v] ? blk1 : vr;
				/* dpalign_row_update_bspoa, bspoa.h:2232-2261 */
				bso_row_shift(a, ur, blk0, q2 - q1, a->smax, a->smin);
				if(q1 == q2){
					if(q1) rh = BSO_SCORE_MIN;
					else if(alnmode == BSO_MODE_OVERLAP || mpos[v] == 0) rh = 0;
					else if(a->pw < 2) rh = O + E * mpos[v];
					else { int c1 = O + E * mpos[v], c2 = Q + P * mpos[v]; rh = c1 > c2 ? c1 : c2; }
				} else if(q1 + bw >= q2) rh = blk0.ub[0];
				else rh = BSO_SCORE_MIN;
				a->mtx = mtx[kprof & 1]; a->hpc = (kprof < 2) ? 1 : 0; /* bspoa.h:2199-2215 */
				bso_row_cal(a, q2, base[v], blk0, dst, rh);
				nupd++;
				if(vst[v]){ bso_row_merge(W, a->pw, blk1, vr, vr); nmrg++; }
				vst[v]++;
				done[v] = 1;
				if((int)vst[v] == nct[v]){
					if(alnmode != BSO_MODE_GLOBAL && q2 + bw >= slen){
						int smax = bso_getscore(a, vr, (int64_t)slen - 1 - q2) + T;
						if(smax > maxscr){ maxscr = smax; maxidx = (int)v; maxoff = (int)slen - 1; }
					}
					stack[sp++] = v;
				}
			}
		}
	}
	#undef NODE_ROW
	best[0] = maxscr; best[1] = maxidx; best[2] = maxoff;
	if(ops){ ops[0] = nupd; ops[1] = nmrg; }
	free(scratch); free(mpos); free(stack); free(vst);
	return a->err;
}

/* Batch runner over the packed arenas of include/bsalign_b200.h (bsb200_poa_rows_batch): a pthread pool over independent
 * sweep jobs.  rows_lin (optional): per job nnode * 3 * bw bytes at rows_off[job] (LINEAR band order); ubs (optional): per
 * node 17 ints; done (optional): per node.  Used by the parity tests and by bench.py's cpu_baseline leg. */
typedef struct {
	uint32_t njobs;
	const int32_t *par; const uint8_t *queries; const uint64_t *qoff; const uint32_t *slen;
	const uint64_t *node_off; const uint8_t *base, *bonus; const int32_t *rpos, *nct, *eoff; const uint64_t *edge_off; const int32_t *edst;
	const uint32_t *head, *tail;
	int8_t *rows_lin; const uint64_t *rows_off; int32_t *ubs; uint8_t *done; int32_t *best; uint64_t *ops;
	volatile uint32_t next; volatile int err;
} bso_poa_batch_t;

static void* bso_poa_worker(void *arg){
	bso_poa_batch_t *b = (bso_poa_batch_t*)arg;
	int err = 0;
	while(1){
		uint32_t i = __sync_fetch_and_add(&b->next, 1);
		uint64_t n0, nn;
		uint32_t bw;
		int8_t *rows; int32_t *ubs; uint8_t *done;
		if(i >= b->njobs) break;
		n0 = b->node_off[i]; nn = b->node_off[i + 1] - n0; bw = (uint32_t)b->par[(size_t)i * 10];
		rows = b->rows_lin ? b->rows_lin + b->rows_off[i] : malloc(nn * 3 * (size_t)bw);
		ubs = b->ubs ? b->ubs + n0 * 17 : malloc(nn * 17 * sizeof(int32_t));
		done = b->done ? b->done + n0 : malloc(nn);
		err |= bso_poa_sweep(b->par + (size_t)i * 10, b->queries + b->qoff[i], b->slen[i], (uint32_t)nn, b->base + n0, b->bonus + n0, b->rpos + n0, b->nct + n0,
			b->eoff + n0 + i, b->edst + b->edge_off[i], b->head[i], b->tail[i], rows, ubs, done, b->best + (size_t)i * 3, b->ops ? b->ops + (size_t)i * 2 : NULL);
		if(!b->rows_lin) free(rows);
		if(!b->ubs) free(ubs);
		if(!b->done) free(done);
	}
	if(err) __sync_fetch_and_or(&b->err, err);
	return NULL;
}

int bso_poa_sweep_batch(uint32_t njobs, const int32_t *par, const uint8_t *queries, const uint64_t *qoff, const uint32_t *slen,
		const uint64_t *node_off, const uint8_t *base, const uint8_t *bonus, const int32_t *rpos, const int32_t *nct,
		const int32_t *eoff, const uint64_t *edge_off, const int32_t *edst, const uint32_t *head, const uint32_t *tail,
		int8_t *rows_lin, const uint64_t *rows_off, int32_t *ubs, uint8_t *done, int32_t *best, uint64_t *ops, int nthreads){
	bso_poa_batch_t b;
	pthread_t *tids;
	int i;
	memset(&b, 0, sizeof(b));
	b.njobs = njobs; b.par = par; b.queries = queries; b.qoff = qoff; b.slen = slen; b.node_off = node_off; b.base = base; b.bonus = bonus;
	b.rpos = rpos; b.nct = nct; b.eoff = eoff; b.edge_off = edge_off; b.edst = edst; b.head = head; b.tail = tail;
	b.rows_lin = rows_lin; b.rows_off = rows_off; b.ubs = ubs; b.done = done; b.best = best; b.ops = ops;
	if(nthreads < 1) nthreads = 1;
	if(nthreads == 1){ bso_poa_worker(&b); return b.err; }
	tids = malloc(sizeof(pthread_t) * nthreads);
	for(i=0;i<nthreads;i++) pthread_create(tids + i, NULL, bso_poa_worker, &b);
	for(i=0;i<nthreads;i++) pthread_join(tids[i], NULL);
	free(tids);
	return b.err;
}

/* ------------------------------------------------------------------------------------------------
 * The walk of alignment2graph_bspoa (bspoa.h:2274-2497): from (maxidx, maxoff) back to the head, re-deriving every step from the
 * node rows.  Only the DECISIONS are restated (which node every read position is matched to, insertions, deletions); the graph
 * surgery the reference interleaves (merge_nodes_bspoa, cpos bookkeeping) stays host code and is replayed from them.
 * Reverse edges: per node the erev list in list order, restricted to selected nodes, with bspoaedge_t.cov.
 * match[x] (x < slen): local node id matched to read position x, or -1.
 * out[8] = {final x (rs.qb before + g->qb), final node, mat, mis, ins, del, start node, flags}.
 * ------------------------------------------------------------------------------------------------ */
#define BSO_SCORE_MAX 536870911   /* SEQALIGN_SCORE_MAX = MAX_B4 >> 2, bsalign.h:59 */
int bso_poa_backtrace(const int32_t *par, const uint8_t *query, uint32_t slen, uint32_t nnode,
		const uint8_t *base, const uint8_t *bonus, const int32_t *rpos,
		const int32_t *reoff, const int32_t *resrc, const int32_t *recov, uint32_t head, uint32_t tail,
		const int8_t *rows, const int32_t *ubs, int32_t midx, int32_t xe, int32_t *match, int32_t *out){
	bso_epi8_t A, *a = &A;
	uint32_t bw = (uint32_t)par[0], W = bw / BSO_LANES, k;
	int alnmode = par[1] & 3, M = par[2], X = par[3], O = par[4], E = par[5], Q = par[6], P = par[7], refbonus = par[9];
	int pw, x = xe, Hs0 = 0, Hs1, Hs2 = 0, bt = -1, mat = 0, mis = 0, ins = 0, del = 0;
	int64_t guard = 0, guard_max = 8 * ((int64_t)slen + nnode) + 64;
	uint32_t n, nidx, urow;   /* urow: the row the reference's `us` pointer last referred to (bspoa.h:2472, 2388) */
	(void)tail;
	memset(a, 0, sizeof(A));
	a->bw = bw; a->W = W; a->qlen = slen;
	pw = bso_epi8_piecewise((int8_t)O, (int8_t)E, (int8_t)Q, (int8_t)P, bw);
	for(k=0;k<slen;k++) match[k] = -1;
	memset(out, 0, 8 * sizeof(int32_t));
	out[6] = midx;
	if(midx < 0 || (uint32_t)midx >= nnode){ out[7] = BSO_ERR_RANGE; out[0] = x; out[1] = midx; return BSO_ERR_RANGE; }
	#define ROW_OF(r, n_) do { (r).u = (int8_t*)rows + (size_t)(n_) * 3 * bw; (r).e = (r).u + bw; (r).q = (r).u + 2 * (size_t)bw; (r).ub = (int32_t*)ubs + (size_t)(n_) * 17; } while(0)
	n = nidx = (uint32_t)midx; urow = n;
	{ bso_row_t r; ROW_OF(r, n); Hs1 = bso_getscore(a, r, (int64_t)x - rpos[n]); }
	while(1){
		if(++guard > guard_max){ a->err |= BSO_ERR_LOOP; break; }
		if(n == head || x < 0) break;
		if(bt == 2 || bt == 4){ /* inside a deletion: leave node n for a predecessor that explains the score (bspoa.h:2308-2356) */
			int32_t ei, found = 0;
			del++;
			for(ei=reoff[n];ei<reoff[n+1];ei++){
				uint32_t w = (uint32_t)resrc[ei];
				bso_row_t r;
				int8_t q;
				if(x < rpos[w] || x >= rpos[w] + (int)bw) continue;
				ROW_OF(r, w); urow = w;
				Hs0 = bso_getscore(a, r, (int64_t)x - rpos[w]);
				if(bt == 2) q = pw ? r.e[x - rpos[w]] : (int8_t)(O + E);
				else q = r.q[x - rpos[w]];
				if(Hs0 + q != Hs1) continue;
				n = w;
				if(q == ((bt == 2) ? O + E : Q + P)){ bt = -1; Hs1 = Hs0; Hs2 = 0; }
				else { Hs1 -= (bt == 2) ? E : P; Hs2++; }
				found = 1;
				break;
			}
			if(!found){ a->err |= BSO_ERR_LOOP; break; } /* the reference would repeat this state forever */
			continue;
		} else if(bt == 1 || bt == 3){ /* insertion run (bspoa.h:2357-2392) */
			int t;
			ins++;
			if(pw == 2){ int c1 = O + E * Hs2, c2 = Q + P * Hs2; t = c1 > c2 ? c1 : c2; } else t = O + E * Hs2;
			x--;
			if(Hs0 + t == Hs1){ bt = -1; Hs1 = Hs0; Hs2 = 0; }
			else if(x >= 0){
				int64_t p = (int64_t)x - rpos[urow];
				if(p < 0 || p >= (int64_t)bw){ a->err |= BSO_ERR_RANGE; break; }
				Hs0 -= rows[(size_t)urow * 3 * bw + p];
				Hs2++;
			}
			continue;
		} else if(bt == 0){ /* match / mismatch of read position x with node n (bspoa.h:2393-2410) */
			match[x] = (int32_t)n;
			if(n != head && n != tail && query[x] == base[n]) mat++; else mis++;
			x--;
			n = nidx;
			bt = -1;
		} else { /* decide the next step from the predecessors of n (bspoa.h:2411-2496) */
			int32_t ei, btc = 0;
			int have = 0, bi = 0, b_h0 = 0; uint32_t b_w = 0;
			int bti_low = 0xFF; /* (bti & 0xFF) of the reference; bti == MAX_U4 initially */
			for(ei=reoff[n];ei<reoff[n+1];ei++){
				uint32_t w = (uint32_t)resrc[ei];
				bso_row_t r;
				int ft = 0, s, scr[3], i, kprof;
				int64_t p;
				ROW_OF(r, w);
				if(x < rpos[w] || x > (int)bw + rpos[w]) continue;
				urow = w;
				if(x == (int)bw + rpos[w]){ Hs0 = bso_getscore(a, r, (int64_t)x - rpos[w] - 1); ft |= (1 << 2) | (1 << 4); }
				else if(x == rpos[w]){
					if(rpos[w] == 0 && (alnmode == BSO_MODE_OVERLAP || w == head)){ Hs0 = r.ub[0]; ft |= 1 << 15; }
					else { Hs0 = r.ub[0]; ft |= 1 << 0; }
				} else Hs0 = bso_getscore(a, r, (int64_t)x - rpos[w] - 1);
				kprof = (base[w] == base[n]) * 2 + bonus[n];
				s = (int8_t)(((query[x] & 3) == base[n]) ? ((kprof & 1) ? M + refbonus : M) : X);
				if(kprof < 2 && (uint32_t)x + 1 < slen && query[x] != query[x + 1]) s += 1; /* the hpc profiles, bsalign.h:2204-2206 */
				if(ft & (1 << 15)) s -= r.ub[0];
				p = (int64_t)x - rpos[w];
				scr[0] = (ft & (1 << 0)) ? BSO_SCORE_MIN : s;
				scr[1] = (ft & (1 << 2)) ? BSO_SCORE_MIN : r.u[p] + (pw ? r.e[p] : E);
				scr[2] = (ft & (1 << 4)) ? BSO_SCORE_MIN : (pw == 2 ? r.u[p] + r.q[p] : BSO_SCORE_MAX);
				for(i=0;i<3;i++){
					if(Hs0 + scr[i] == Hs1){
						if(recov[ei] > btc){ have = 1; bi = i; bti_low = i; b_w = w; b_h0 = Hs0; btc = recov[ei]; }
						else if(recov[ei] == btc && i == 0 && bti_low != 0){ have = 1; bi = 0; bti_low = 0; b_w = w; b_h0 = Hs0; btc = recov[ei]; }
					}
				}
			}
			if(!have){
				int64_t p = (int64_t)x - rpos[n];
				if(p < 0 || p >= (int64_t)bw){ a->err |= BSO_ERR_RANGE; break; }
				bt = 1; Hs2 = 1; urow = n;
				Hs0 = Hs1 - rows[(size_t)n * 3 * bw + p];
			} else if(bi == 0){ bt = 0; nidx = b_w; Hs1 = b_h0; Hs2 = 0; }
			else if(bi == 1){ bt = 2; Hs2 = 1; }
			else { bt = 4; Hs2 = 1; }
		}
	}
	#undef ROW_OF
	out[0] = x; out[1] = (int32_t)n; out[2] = mat; out[3] = mis; out[4] = ins; out[5] = del; out[7] = a->err;
	return a->err;
}

/* ------------------------------------------------------------------------------------------------
 * k-mer guided edit alignment: kmer_striped_seqedit_pairwise, bsalign.h:1209-1536 (main.c:196 `edit -m kmer -k ksz`,
 * bspoa.h:2089 band placement of long reads).  Restated as: unique shared canonical k-mers (:1230-1276), the reference's chain over
 * them (:1277-1334; note its predecessor rule for a replaced tail), the offset filter (:1347-1394), the coverage tests, and the
 * stitching of per-gap edit alignments around the anchors (:1446-1533) with its cigar-push order kept as it is.
 * ------------------------------------------------------------------------------------------------ */
typedef struct { uint32_t kmer, src, dir, off; } bso_km_t;
typedef struct { uint32_t q, t, on; } bso_hit_t;

static int bso_km_cmp(const void *a, const void *b){
	const bso_km_t *x = (const bso_km_t*)a, *y = (const bso_km_t*)b;
	if(x->kmer != y->kmer) return x->kmer < y->kmer ? -1 : 1;
	return (int)x->src - (int)y->src;
}
static int bso_hit_cmp(const void *a, const void *b){
	const bso_hit_t *x = (const bso_hit_t*)a, *y = (const bso_hit_t*)b;
	return x->q < y->q ? -1 : (x->q > y->q);
}
static int bso_int_cmp(const void *a, const void *b){ int x = *(const int*)a, y = *(const int*)b; return x < y ? -1 : (x > y); }

/* anchors of a pair in query order; returns their number (0 = the caller falls back to the plain global edit) */
uint32_t bso_kmer_anchors(uint32_t ksz, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen, uint32_t *aq, uint32_t *at){
	uint32_t cmin, kmk, sft, nk = 0, nh = 0, i, b, e, m, len, kept = 0, fw, rv;
	bso_km_t *km;
	bso_hit_t *hit;
	uint32_t *tails, *pred;
	int *dl;
	if(ksz > 15) ksz = 15;
	if(ksz == 0) return 0;
	cmin = (uint32_t)((qlen < tlen ? qlen : tlen) * 0.05 + 1); /* :1221-1222 */
	if(cmin > 2 * ksz) cmin = 2 * ksz;
	kmk = 0xFFFFFFFFu >> ((16 - ksz) << 1);
	sft = (ksz - 1) << 1;
	km = malloc(sizeof(bso_km_t) * ((size_t)qlen + tlen + 1));
	for(int src=0;src<2;src++){ /* canonical k-mers of both sequences (:1236-1255) */
		const uint8_t *s = src ? tseq : qseq;
		uint32_t n = src ? tlen : qlen;
		if(n < ksz) continue;
		fw = rv = 0;
		for(i=0;i<n;i++){
			uint32_t c = s[i];
			fw = ((fw << 2) | c) & kmk;
			rv = (rv >> 2) | (((~c) & 3u) << sft);
			if(i + 1 >= ksz){
				uint32_t d = rv < fw;
				km[nk].kmer = (d ? rv : fw) & 0x3FFFFFFFu; km[nk].src = src; km[nk].dir = d; km[nk].off = i + 1 - ksz; nk++;
			}
		}
	}
	qsort(km, nk, sizeof(bso_km_t), bso_km_cmp);
	hit = malloc(sizeof(bso_hit_t) * ((size_t)(qlen < tlen ? qlen : tlen) + 1));
	/* k-mers seen exactly twice, once per sequence, same strand (:1259-1273).  The reference ends its scan on a zeroed sentinel, so a
	 * final group of k-mer value 0 is never looked at. */
	for(b=0;b<nk;b=e){
		for(e=b+1;e<nk&&km[e].kmer==km[b].kmer;e++);
		if(e == nk && km[b].kmer == 0) break;
		if(e - b == 2 && km[b].src != km[b + 1].src && km[b].dir == km[b + 1].dir){ hit[nh].q = km[b].off; hit[nh].t = km[b + 1].off; hit[nh].on = 0; nh++; }
	}
	free(km);
	if(nh * ksz < cmin){ free(hit); return 0; }
	qsort(hit, nh, sizeof(bso_hit_t), bso_hit_cmp);
	/* chain (:1281-1330): patience tails; an element that replaces an inner tail inherits the predecessor OF THE TAIL BEFORE IT */
	tails = malloc(sizeof(uint32_t) * nh); pred = malloc(sizeof(uint32_t) * nh);
	tails[0] = 0; pred[0] = 0xFFFFFFFFu; len = 1;
	for(i=1;i<nh;i++){
		if(hit[i].t > hit[tails[len - 1]].t){ pred[i] = tails[len - 1]; tails[len++] = i; }
		else if(hit[i].t <= hit[tails[0]].t){ pred[i] = 0xFFFFFFFFu; tails[0] = i; }
		else {
			b = 0; e = len;
			while(b < e){
				m = b + ((e - b) >> 1);
				if(hit[i].t > hit[tails[m]].t) b = m + 1;
				else if(hit[i].t < hit[tails[m]].t) e = m;
				else { b = m; break; }
			}
			pred[i] = pred[tails[b - 1]];
			tails[b] = i;
		}
	}
	b = 0; e = 0xFFFFFFFFu;
	for(m=tails[len-1];m!=0xFFFFFFFFu;m=pred[m]){
		hit[m].on = 1;
		if(hit[m].t + ksz <= e) b += ksz; else b += e - hit[m].t;
		e = hit[m].t;
	}
	free(tails); free(pred);
	if(b < cmin){ free(hit); return 0; }
	/* drop anchors whose diagonal is far from the mean (:1347-1394) */
	dl = malloc(sizeof(int) * nh);
	while(1){
		int tot = 0, mean, median, var, d;
		uint32_t cnt = 0, drop = 0;
		for(i=0;i<nh;i++) if(hit[i].on){ d = (int)hit[i].q - (int)hit[i].t; tot += d; dl[cnt++] = d; }
		if(cnt * ksz < cmin) break;
		mean = tot / (int)cnt;
		qsort(dl, cnt, sizeof(int), bso_int_cmp);
		median = dl[cnt / 2]; /* quick_median_array, sort.h:268-310: the element of rank size/2 */
		var = (median > mean ? median - mean : mean - median) * 3;
		if(var < 50) var = 50;
		for(i=0;i<nh;i++) if(hit[i].on){
			d = (int)hit[i].q - (int)hit[i].t - mean;
			if((d < 0 ? -d : d) > var){ hit[i].on = 0; drop++; }
		}
		if(drop == 0) break;
	}
	free(dl);
	m = 0; e = 0;
	for(i=0;i<nh;i++) if(hit[i].on){ /* target bases covered by the kept anchors (:1403-1413) */
		if(hit[i].t >= e + ksz) m += ksz; else m += hit[i].t + ksz - e;
		e = hit[i].t + ksz;
		aq[kept] = hit[i].q; at[kept] = hit[i].t; kept++;
	}
	free(hit);
	if(m < cmin) return 0;
	return kept;
}

typedef struct { uint32_t *buf; uint32_t cap, n; int over; } bso_cv_t;
static void bso_cv_raw(bso_cv_t *v, uint32_t w){ if(v->buf && v->n < v->cap) v->buf[v->n] = w; else if(v->buf) v->over = 1; v->n++; }
static void bso_cv_push(bso_cv_t *v, uint32_t op, uint32_t sz){ /* _push_cigar_u4v, :401-407 */
	if(v->n && v->buf && v->n <= v->cap && (v->buf[v->n - 1] & 0xf) == op) v->buf[v->n - 1] += sz << 4;
	else bso_cv_raw(v, sz << 4 | op);
}

/* cigar must have room for 2 * (qlen + tlen) + 4 words */
int bso_kmer_edit_pairwise(uint32_t ksz, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen,
		int32_t *rs, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar){
	uint32_t *aq = malloc(sizeof(uint32_t) * ((size_t)qlen + 1)), *at = malloc(sizeof(uint32_t) * ((size_t)qlen + 1));
	uint32_t kmap, i, k, qb = 0, tb = 0, qe, te, ml = 0, mode, nseg;
	int32_t r2[10];
	int err = 0;
	bso_cv_t cv;
	if(ksz > 15) ksz = 15;
	kmap = bso_kmer_anchors(ksz, qseq, qlen, tseq, tlen, aq, at);
	if(kmap == 0){
		free(aq); free(at);
		return bso_edit_pairwise(qseq, qlen, tseq, tlen, BSO_MODE_GLOBAL, 0, rs, cigar, cigar_cap, ncigar);
	}
	memset(rs, 0, 10 * sizeof(int32_t));
	cv.buf = cigar; cv.cap = cigar_cap; cv.n = 0; cv.over = 0;
	mode = 3;
	for(i=0;i<=kmap;i++){
		if(i == kmap){ qe = qlen; te = tlen; mode = BSO_MODE_EXTEND; }
		else { qe = aq[i] + ksz / 2; te = at[i] + ksz / 2; ml++; }
		if(!(qb == qe && tb == te)){
			uint32_t sq = qe - qb, st = te - tb, *seg = malloc(sizeof(uint32_t) * ((size_t)sq + st + 2));
			if(ml){ bso_cv_push(&cv, 0, ml); rs[5] += ml; rs[9] += ml; ml = 0; }
			if(mode == 3){ /* the stretch before the first anchor: EXTEND on the reversed prefixes, then everything so far is reversed (:1489-1499) */
				uint8_t *rq = malloc(sq + 1), *rt = malloc(st + 1);
				for(k=0;k<sq;k++) rq[k] = qseq[qe - 1 - k];
				for(k=0;k<st;k++) rt[k] = tseq[te - 1 - k];
				err |= bso_edit_pairwise(rq, sq, rt, st, BSO_MODE_EXTEND, 0, r2, seg, sq + st + 2, &nseg);
				free(rq); free(rt);
				for(k=0;k<nseg;k++) bso_cv_raw(&cv, seg[k]);
				rs[1] = qe - r2[2]; rs[3] = te - r2[4]; rs[2] = qe; rs[4] = te;
				if(cv.buf) for(k=0;k<cv.n/2&&cv.n<=cv.cap;k++){ uint32_t w = cv.buf[k]; cv.buf[k] = cv.buf[cv.n - 1 - k]; cv.buf[cv.n - 1 - k] = w; }
			} else {
				err |= bso_edit_pairwise(qseq + qb, sq, tseq + tb, st, (int)mode, 0, r2, seg, sq + st + 2, &nseg);
				for(k=0;k<nseg;k++) bso_cv_raw(&cv, seg[k]);
				rs[2] = qb + r2[2]; rs[4] = tb + r2[4];
			}
			free(seg);
			rs[5] += r2[5]; rs[6] += r2[6]; rs[7] += r2[7]; rs[8] += r2[8]; rs[9] += r2[9]; rs[0] += r2[0];
		}
		qb = qe + 1; tb = te + 1;
		mode = BSO_MODE_GLOBAL;
	}
	if(ncigar) *ncigar = cv.n;
	if(cv.over) err |= BSO_ERR_CIGCAP;
	free(aq); free(at);
	return err;
}

/* ------------------------------------------------------------------------------------------------
 * Re-alignment of one read against the MSA profile: remsa_pedit_rd_bspoacore, bspoa.h:3916-4045, restated
 * per cell.  An anti-diagonal band of bw cells runs down the main diagonal of the (read-in-MSA-coordinates) x (MSA columns) square from
 * mbeg to mend; diagonal `moff` = x + y is row moff + 1 of two byte matrices of bw + 2 bytes per row (one border byte on each side) that hold
 * the DP in difference form: U = H - (value from above), V = H - (value from the left), unsigned bytes with saturating arithmetic
 * (maxmat_dp_diag_rowcal, bspoa.h:3856-3896).  The cell score is the profile count of the read's base in that column plus the read's
 * homopolymer count for the consensus base (prepare: bspoa.h:3760-3784; seqs1 / mats1 are stored reversed so that a diagonal is
 * contiguous).  The walk back (bspoa.h:3962-4040) re-derives every step from the differences and reports, per read position, the MSA
 * column it is matched to (the reference merges the read's node into that column's node there).
 * seqs0 / seqs1 / mats point at index 0 of arrays with bw / 2 readable bytes in front and behind.  M0 / M1: (2 * mlen + 1) * (bw + 2) bytes.
 * ------------------------------------------------------------------------------------------------ */
static inline int bso_remsa_score(const uint8_t *seqs0, const uint8_t *seqs1, const uint8_t *const mats0[4], const uint8_t *const mats1[4], int mlen, int xi, int yi){
	const int s1 = seqs1[mlen - 1 - yi], s0 = seqs0[xi];
	int h = (s1 < 4 ? mats0[s1][xi] : 0) + (s0 < 4 ? mats1[s0][mlen - 1 - yi] : 0);
	return h > 255 ? 255 : h;
}

int bso_remsa_core(int mlen, int bw, int mbeg, int mend, int rend, const uint8_t *seqs0, const uint8_t *seqs1,
		const uint8_t *m00, const uint8_t *m01, const uint8_t *m02, const uint8_t *m03, const uint8_t *m10, const uint8_t *m11, const uint8_t *m12, const uint8_t *m13,
		uint8_t *M0, uint8_t *M1, int32_t *match, int32_t *scr_out){
	const uint8_t *const mats0[4] = {m00, m01, m02, m03}, *const mats1[4] = {m10, m11, m12, m13};
	const int rowlen = bw + 2, half = bw / 2;
	int x, y, i, c, dir, xi, yi, roff, scr = 0, err = 0;
	for(c=0;c<rend;c++) match[c] = -1;
	/* init (bspoa.h:3749-3758) */
	memset(M0 + (size_t)rowlen * 2 * mbeg, 0, rowlen); memset(M1 + (size_t)rowlen * 2 * mbeg, 0, rowlen);
	M0[(size_t)rowlen * 2 * mbeg + 1 + half - 1] = 255; M1[(size_t)rowlen * 2 * mbeg + 1 + half] = 255;
	x = y = mbeg;
	for(i=x+y;;i++){
		const uint8_t *pu = M0 + (size_t)rowlen * i + 1, *pv = M1 + (size_t)rowlen * i + 1;
		uint8_t *nu = M0 + (size_t)rowlen * (i + 1) + 1, *nv = M1 + (size_t)rowlen * (i + 1) + 1;
		dir = i & 1;
		/* on this diagonal cell c is (x - half + c, y + half - c): both x and y of the band centre are the loop's (x, y) */
		for(c=0;c<bw;c++){
			int h = bso_remsa_score(seqs0, seqs1, mats0, mats1, mlen, x - half + c, y + half - c);
			const int u = dir ? pu[c + 1] : pu[c], v = dir ? pv[c] : pv[c - 1];
			if(h < u) h = u;
			if(h < v) h = v;
			nu[c] = (uint8_t)(h - v); nv[c] = (uint8_t)(h - u);   /* h >= u, v: the saturating subtraction never saturates */
		}
		if(dir){ nu[-1] = 255; nv[-1] = 0; nu[bw] = 0; nv[bw] = 0; }
		else { nu[-1] = 0; nv[-1] = 0; nu[bw] = 0; nv[bw] = 255; }
		if(dir) y++; else x++;
		if(x >= mend) break;
	}
	/* the walk (bspoa.h:3962-4040) */
	xi = yi = mend - 1; roff = rend;
	while(xi >= 0 && yi >= 0){
		int xx, h, e, f, s, mdir;
		i = xi + yi;
		if(i < mbeg + mbeg) break;
		dir = mdir = i & 1;
		xx = (xi - yi - mdir) / 2 + half;   /* C division, as in the reference */
		if(xx < 0 || xx >= bw){ err |= BSO_ERR_RANGE; break; }
		{
			const uint8_t *pu = M0 + (size_t)rowlen * i + 1, *pv = M1 + (size_t)rowlen * i + 1, *nu = M0 + (size_t)rowlen * (i + 1) + 1;
			h = bso_remsa_score(seqs0, seqs1, mats0, mats1, mlen, xi, yi);
			if(dir){ e = pu[xx + 1]; f = pv[xx]; } else { e = pu[xx]; f = pv[xx - 1]; }
			s = f + nu[xx];
		}
		if(s == f && !(xx == 0 && dir == 0)){ if(seqs0[xi] < 4) roff--; xi--; }
		else if(s == e){ yi--; }
		else if(s == h){
			if(seqs0[xi] < 4){ roff--; if(roff >= 0 && roff < rend) match[roff] = yi; else err |= BSO_ERR_RANGE; }
			scr += s; xi--; yi--;
		} else { err |= BSO_ERR_LOOP; break; }
	}
	if(scr_out) *scr_out = scr;
	return err;
}
