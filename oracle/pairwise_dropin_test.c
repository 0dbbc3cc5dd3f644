/*
 * oracle/pairwise_dropin_test.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Drop-in check of include/bsalign_b200_compat.h against the UNMODIFIED reference headers (compiled from where they lie under
 * /root/reference, see oracle/Makefile): the README program of the reference (README.md:51-80) with every pair going through
 *   (a) banded_striped_epi8_seqalign_pairwise / striped_seqedit_pairwise / kmer_striped_seqedit_pairwise of the reference (bsalign.h:3854, :1046, :1209), and
 *   (b) the same NAMES re-bodied on the GPU (b200_banded_striped_epi8_seqalign_pairwise / b200_striped_seqedit_pairwise / b200_kmer_striped_seqedit_pairwise),
 * same mempool / cigars vectors, modes with and without SEQALIGN_MODE_CIGRESV; seqalign_result_t and every cigar word must be equal.
 * Also checks that a pair the library flags (the reference's traceback never terminates on it) does not come back looking valid:
 * the status hook is called.
 * Usage: pairwise_dropin <pairs> <qlen> <seed>
 */
static int dropin_flagged = 0;
#define BSALIGN_B200_ON_STATUS(st, func) do { dropin_flagged ++; } while(0)
#include "bsalign.h"
#include "bsalign_b200_compat.h"

static uint64_t rng_state;
static inline uint64_t rng_next(void){ rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27; return rng_state * 2685821657736338717ULL; }
static inline double rng_unif(void){ return (rng_next() >> 11) * (1.0 / 9007199254740992.0); }

static u4i mutate(const u1i *tmpl, u4i tlen, u1i *out, double ps, double pi, double pd){
	u4i i, n = 0;
	for(i=0;i<tlen;i++){
		double r = rng_unif();
		if(r < ps){ out[n ++] = (tmpl[i] + 1 + (rng_next() % 3)) & 3; }
		else if(r < ps + pi){ out[n ++] = tmpl[i]; out[n ++] = rng_next() & 3; }
		else if(r < ps + pi + pd){ }
		else out[n ++] = tmpl[i];
	}
	return n;
}

static int same_result(seqalign_result_t *a, seqalign_result_t *b){ return memcmp(a, b, sizeof(seqalign_result_t)) == 0; }
static int same_cigars(u4v *a, u4v *b){ return a->size == b->size && (a->size == 0 || memcmp(a->buffer, b->buffer, a->size * sizeof(u4i)) == 0); }

int main(int argc, char **argv){
	u4i npairs = argc > 1 ? atoi(argv[1]) : 24, qlen = argc > 2 ? atoi(argv[2]) : 500, p, bad = 0, done = 0;
	b1i mtx[16];
	b1v *mempool = adv_init_b1v(1024, 0, WORDSIZE, 0);
	u4v *ca = init_u4v(64), *cb = init_u4v(64);
	u1i *q = malloc(qlen + 16), *t = malloc(2 * (size_t)qlen + 16);
	seqalign_result_t ra, rb;
	static const int modes[3] = {SEQALIGN_MODE_GLOBAL, SEQALIGN_MODE_OVERLAP, SEQALIGN_MODE_EXTEND};
	static const u4i bands[4] = {0, 64, 128, 48};
	rng_state = (argc > 3 ? strtoull(argv[3], NULL, 10) : 1) * 0x9E3779B97F4A7C15ULL + 88172645463325252ULL;
	banded_striped_epi8_seqalign_set_score_matrix(mtx, 2, -6);
	for(p=0;p<npairs;p++){
		u4i i, tlen, ql = qlen - (p % 7) * 11;
		int mode = modes[p % 3], resv = (p % 5 == 4) ? SEQALIGN_MODE_CIGRESV : 0;
		u4i bw = bands[p % 4];
		for(i=0;i<ql;i++) q[i] = rng_next() & 3;
		tlen = mutate(q, ql, t, 0.03, 0.03, 0.04);
		if(tlen == 0){ t[0] = q[0]; tlen = 1; }
		
		/* 8-bit affine and two-piece */
		clear_u4v(ca); clear_u4v(cb);
		if(resv){ push_u4v(ca, 0x77); push_u4v(cb, 0x77); }
		ra = banded_striped_epi8_seqalign_pairwise(q, ql, t, tlen, mempool, ca, mode | resv, bw, mtx, -3, -2, (p & 1) ? -8 : 0, (p & 1) ? -1 : 0, 0);
		rb = b200_banded_striped_epi8_seqalign_pairwise(q, ql, t, tlen, mempool, cb, mode | resv, bw, mtx, -3, -2, (p & 1) ? -8 : 0, (p & 1) ? -1 : 0, 0);
		if(!same_result(&ra, &rb) || !same_cigars(ca, cb)){ bad ++; fprintf(stderr, "pair %u epi8 mode %d bw %u: score %d vs %d, cigars %u vs %u\n", p, mode, bw, ra.score, rb.score, (u4i)ca->size, (u4i)cb->size); }
		done ++;
		/* 2-bit edit */
		clear_u4v(ca); clear_u4v(cb);
		ra = striped_seqedit_pairwise(q, ql, t, tlen, mode, (mode == SEQALIGN_MODE_GLOBAL) ? 64 : 0, mempool, ca, 0);
		rb = b200_striped_seqedit_pairwise(q, ql, t, tlen, mode, (mode == SEQALIGN_MODE_GLOBAL) ? 64 : 0, mempool, cb, 0);
		if(!same_result(&ra, &rb) || !same_cigars(ca, cb)){ bad ++; fprintf(stderr, "pair %u edit mode %d: score %d vs %d, cigars %u vs %u\n", p, mode, ra.score, rb.score, (u4i)ca->size, (u4i)cb->size); }
		done ++;
		/* k-mer guided edit (bsalign.h:1209): it reverses a prefix of both sequences in place and restores it */
		clear_u4v(ca); clear_u4v(cb); push_u4v(cb, 0x55);
		ra = kmer_striped_seqedit_pairwise((p & 1) ? 13 : 7, q, ql, t, tlen, mempool, ca, 0);
		rb = b200_kmer_striped_seqedit_pairwise((p & 1) ? 13 : 7, q, ql, t, tlen, mempool, cb, 0);
		if(!same_result(&ra, &rb) || !same_cigars(ca, cb)){ bad ++; fprintf(stderr, "pair %u kmer edit: score %d vs %d, cigars %u vs %u\n", p, ra.score, rb.score, (u4i)ca->size, (u4i)cb->size); }
		done ++;
	}
	/* the empty-input rule of the edit entry point (bsalign.h:1051-1054) */
	rb = b200_striped_seqedit_pairwise(q, 0, t, 5, SEQALIGN_MODE_GLOBAL, 0, mempool, cb, 0);
	memset(&ra, 0, sizeof(ra));
	if(!same_result(&ra, &rb)){ bad ++; fprintf(stderr, "empty edit input: result not all zero\n"); }
	/* pairs the reference cannot trace back: with scores that saturate int8 (M 30, X -40, O -40, E -20) its re-derived traceback finds
	   no consistent predecessor and never leaves its insertion search (bsalign.h:3798-3814).  Only the GPU entry is called; it must
	   report such a pair through the status hook instead of returning a valid-looking result */
	{
		u4i i, before = dropin_flagged;
		b1i sat[16];
		banded_striped_epi8_seqalign_set_score_matrix(sat, 30, -40);
		for(i=0;i<8;i++){
			u4i k, tl;
			for(k=0;k<440;k++) q[k] = rng_next() & 3;
			tl = mutate(q, 440, t, 0.07, 0.06, 0.07);
			rb = b200_banded_striped_epi8_seqalign_pairwise(q, 440, t, tl, mempool, cb, SEQALIGN_MODE_GLOBAL, 48, sat, -40, -20, 0, 0, 0);
		}
		if(dropin_flagged == (int)before){ bad ++; fprintf(stderr, "8 pairs under int8-saturating scores: none was flagged\n"); }
	}
	printf("pairwise_dropin: calls=%u identical=%u/%u flagged_calls=%d last_status=%d\n", done, done - bad, done, dropin_flagged, bsalign_b200_last_status);
	free(q); free(t); free_u4v(ca); free_u4v(cb);
	return bad ? 1 : 0;
}
