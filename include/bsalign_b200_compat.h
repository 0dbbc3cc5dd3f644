/*
 * bsalign_b200_compat.h -- GNU-C drop-in for the two pairwise entry points of ruanjue/bsalign, re-bodied on
 * top of libbsalign_b200.so (C ABI: bsalign_b200.h).
 *
 * Usage in a program that already uses the reference library (README.md:51-80 of the reference):
 *
 *     #include "bsalign.h"                 // the reference's own header: types, cigar helpers, printing
 *     #include "bsalign_b200_compat.h"     // after it
 *     ...
 *     rs = b200_banded_striped_epi8_seqalign_pairwise(qseq, qlen, tseq, tlen, mempool, cigars,
 *                                                     mode, bandwidth, matrix, gapo1, gape1, gapo2, gape2, verbose);
 *
 * or compile with -DBSALIGN_B200_OVERRIDE to have the reference NAMES themselves routed to the GPU
 * (the macros below rename every later use of the two functions).  Signatures, the by-value
 * seqalign_result_t, the u4v cigar vector (cleared, or appended to under SEQALIGN_MODE_CIGRESV) and the
 * "all-zero result for an empty edit input" rule are the reference's (bsalign.h:399, :232, :3713-3719,
 * :1051-1054).  `mempool` is accepted and ignored: scratch lives in HBM.  `verbose` is ignored.
 * Programmer errors keep the reference's behaviour: message on stderr + abort().
 *
 * One pair per call pays a host<->device round trip; throughput needs the batch entry points of
 * bsalign_b200.h (see INTEGRATION.md for the main.c loop rewritten on top of them).
 */
#ifndef BSALIGN_B200_COMPAT_H
#define BSALIGN_B200_COMPAT_H

#include "bsalign_b200.h"

#ifndef BAND_STRIPED_DNA_SEQ_ALIGNMENT_RJ_H
#error "include the reference's bsalign.h before bsalign_b200_compat.h"
#endif

/* one context for the whole program (the library owns it, so every translation unit sees the same one) */
static inline bsb200_ctx *bsalign_b200_default_ctx(void){
	bsb200_ctx *ctx = bsb200_default_context();
	if(ctx == NULL){
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: no CUDA device, and there is no CPU fallback in %s -- %s:%d --\n", __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
		abort();
	}
	return ctx;
}

/*
 * Per-pair status of the last call (bsalign_b200.h: BSB200_ST_*).  The reference has no error codes; on the inputs the library
 * flags it either reads outside its traceback band / never leaves its traceback loop (RANGE, LOOP: a global alignment forced
 * through a band that cannot hold it, bsalign.h:3199-3202, :3798-3814) or computes on stale scratch words (REFBUG, bsalign.h:704-713).
 * A flagged result must not look valid to the caller: RANGE and LOOP end the program like the reference's own programmer errors
 * (message + abort()), REFBUG is reported on stderr and the result (which follows the intended algorithm) is returned.
 * Define BSALIGN_B200_ON_STATUS(st, func) before including this header to handle the flags yourself.
 */
static int bsalign_b200_last_status = 0;
#ifndef BSALIGN_B200_ON_STATUS
#define BSALIGN_B200_ON_STATUS(st, func) do { \
	if((st) & (BSB200_ST_RANGE | BSB200_ST_LOOP)){ \
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: status %d in %s: the reference's traceback leaves its band / never terminates on this pair (bsalign.h:3199, :3798) -- %s:%d --\n", (int)(st), func, __FILE__, __LINE__); fflush(stderr); \
		abort(); \
	} else if((st) & BSB200_ST_REFBUG){ \
		fprintf(stderr, " -- bsalign_b200: status %d in %s: the reference shifts this edit band through stale scratch words (bsalign.h:704-713); result follows the intended algorithm --\n", (int)(st), func); \
	} } while(0)
#endif

static inline void bsalign_b200_take_cigars(u4v *cigars, int mode, const uint32_t *buf, uint32_t n){
	uint32_t i;
	if(cigars == NULL) return;
	if(!(mode & SEQALIGN_MODE_CIGRESV)) clear_u4v(cigars);
	for(i=0;i<n;i++) push_u4v(cigars, buf[i]);
}

static inline seqalign_result_t b200_banded_striped_epi8_seqalign_pairwise(u1i *qseq, u4i qlen, u1i *tseq, u4i tlen, b1v *mempool, u4v *cigars,
		int mode, u4i bandwidth, b1i matrix[16], b1i gapo1, b1i gape1, b1i gapo2, b1i gape2, int verbose){
	seqalign_result_t rs;
	bsb200_result_t r;
	uint32_t *buf, n = 0;
	int32_t st = 0;
	UNUSED(mempool); UNUSED(verbose);
	buf = cigars? (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)qlen + tlen + 2)) : NULL;
	if(bsb200_epi8_pairwise(bsalign_b200_default_ctx(), qseq, qlen, tseq, tlen, seqalign_mode_type(mode), bandwidth, (const int8_t*)matrix,
			gapo1, gape1, gapo2, gape2, &r, buf, qlen + tlen + 2, &n, &st)){
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: %s in %s -- %s:%d --\n", bsb200_last_error(bsalign_b200_default_ctx()), __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
		abort();
	}
	bsalign_b200_last_status = st;
	if(st & ~BSB200_ST_EMPTY) BSALIGN_B200_ON_STATUS(st, __FUNCTION__);
	rs.score = r.score; rs.qb = r.qb; rs.qe = r.qe; rs.tb = r.tb; rs.te = r.te;
	rs.mat = r.mat; rs.mis = r.mis; rs.ins = r.ins; rs.del = r.del; rs.aln = r.aln;
	bsalign_b200_take_cigars(cigars, mode, buf, n);
	if(buf) free(buf);
	return rs;
}

static inline seqalign_result_t b200_striped_seqedit_pairwise(u1i *qseq, u4i qlen, u1i *tseq, u4i tlen, int mode, u4i bandwidth, b1v *mempool, u4v *cigars, int verbose){
	seqalign_result_t rs;
	bsb200_result_t r;
	uint32_t *buf, n = 0;
	int32_t st = 0;
	UNUSED(mempool); UNUSED(verbose);
	if(qlen == 0 || tlen == 0){ memset(&rs, 0, sizeof(seqalign_result_t)); return rs; } /* bsalign.h:1051-1054 */
	buf = cigars? (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)qlen + tlen + 2)) : NULL;
	if(bsb200_edit_pairwise(bsalign_b200_default_ctx(), qseq, qlen, tseq, tlen, seqalign_mode_type(mode), bandwidth, &r, buf, qlen + tlen + 2, &n, &st)){
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: %s in %s -- %s:%d --\n", bsb200_last_error(bsalign_b200_default_ctx()), __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
		abort();
	}
	bsalign_b200_last_status = st;
	if(st & ~BSB200_ST_EMPTY) BSALIGN_B200_ON_STATUS(st, __FUNCTION__);
	rs.score = r.score; rs.qb = r.qb; rs.qe = r.qe; rs.tb = r.tb; rs.te = r.te;
	rs.mat = r.mat; rs.mis = r.mis; rs.ins = r.ins; rs.del = r.del; rs.aln = r.aln;
	bsalign_b200_take_cigars(cigars, mode, buf, n);
	if(buf) free(buf);
	return rs;
}

/* kmer_striped_seqedit_pairwise, bsalign.h:1209: the cigars are always cleared first (bsalign.h:1438), an empty sequence gives the all-zero result */
static inline seqalign_result_t b200_kmer_striped_seqedit_pairwise(u1i ksz, u1i *qseq, u4i qlen, u1i *tseq, u4i tlen, b1v *mempool, u4v *cigars, int verbose){
	seqalign_result_t rs;
	bsb200_result_t r;
	uint32_t *buf, n = 0;
	int32_t st = 0;
	UNUSED(mempool); UNUSED(verbose);
	if(cigars) clear_u4v(cigars);
	if(qlen == 0 || tlen == 0){ memset(&rs, 0, sizeof(seqalign_result_t)); return rs; }
	buf = cigars? (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)qlen + tlen + 2)) : NULL;
	if(bsb200_kmer_edit_pairwise(bsalign_b200_default_ctx(), ksz, qseq, qlen, tseq, tlen, &r, buf, qlen + tlen + 2, &n, &st)){
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: %s in %s -- %s:%d --\n", bsb200_last_error(bsalign_b200_default_ctx()), __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
		abort();
	}
	bsalign_b200_last_status = st;
	if(st & ~BSB200_ST_EMPTY) BSALIGN_B200_ON_STATUS(st, __FUNCTION__);
	rs.score = r.score; rs.qb = r.qb; rs.qe = r.qe; rs.tb = r.tb; rs.te = r.te;
	rs.mat = r.mat; rs.mis = r.mis; rs.ins = r.ins; rs.del = r.del; rs.aln = r.aln;
	bsalign_b200_take_cigars(cigars, 0, buf, n);
	if(buf) free(buf);
	return rs;
}

#ifdef BSALIGN_B200_OVERRIDE
#define banded_striped_epi8_seqalign_pairwise b200_banded_striped_epi8_seqalign_pairwise
#define striped_seqedit_pairwise b200_striped_seqedit_pairwise
#define kmer_striped_seqedit_pairwise b200_kmer_striped_seqedit_pairwise
#endif

#endif
