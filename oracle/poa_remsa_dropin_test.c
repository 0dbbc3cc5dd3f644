/*
 * oracle/poa_remsa_dropin_test.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Drop-in check of the re-alignment driver of include/bsalign_b200_poa_compat.h (b200_poa_realign_run: remsa_pedits_bspoa,
 * bspoa.h:4178-4457, of all in-flight objects with the DP of remsa_pedit_rd_bspoacore, bspoa.h:3916-4045, as batches) against the
 * reference.  Built by oracle/Makefile (target dropin_remsa) against a TEMPORARY copy of the reference's bspoa.h with the one-line
 * change include/bsalign_b200_poa_remsa.h documents (the callee's name at bspoa.h:4451); every other line is the reference's own.
 *   arm A: the reference's end_bspoa (bspoa.h:4722) - in this process no thread of arm A has the hook installed, so B200_REMSA_CORE
 *          calls remsa_pedit_rd_bspoacore;
 *   arm B, default build: b200_end_bspoa_batch - sweeps, walks, band placement AND the re-alignment DPs on the GPU;
 *   arm B, -DREMSA_CPU_CHECK: the reference's end_bspoa with realn = 0 builds the graphs, then b200_poa_realign_batch runs the realn
 *          rounds with the batch entry point replaced by the oracle's scalar core: exercises the rendezvous, the hook and the merge
 *          replay on a box without a GPU (tests/test_poa.py).
 * Consensus, qualities, alternative bases and the whole MSA matrix must be byte-identical.
 * Usage: poa_remsa_dropin <jobs> <reads per job> <template length> <seed> [realn] [reads per job vary by up to this many] [b200_poa_remsa_min_objects]
 */
#include "bsalign.h"
#include "bsalign_b200_poa_kmer.h"
#include "bsalign_b200_poa_remsa.h"
#ifdef REMSA_CPU_CHECK
#include "bsalign_b200.h"
int bso_remsa_core(int mlen, int bw, int mbeg, int mend, int rend, const uint8_t *seqs0, const uint8_t *seqs1,
		const uint8_t *m00, const uint8_t *m01, const uint8_t *m02, const uint8_t *m03, const uint8_t *m10, const uint8_t *m11, const uint8_t *m12, const uint8_t *m13,
		uint8_t *M0, uint8_t *M1, int32_t *match, int32_t *scr_out);
static int cpu_remsa_batch(bsb200_ctx *ctx, uint32_t n, const int32_t *hdr, const uint8_t *in, const uint64_t *in_off, uint64_t in_bytes,
		int32_t *match, const uint64_t *match_off, uint64_t match_ints, int32_t *out, uint8_t *matrices, const uint64_t *mat_off){
	uint32_t j;
	(void)ctx; (void)in_bytes; (void)match_ints; (void)matrices; (void)mat_off;
	for(j=0;j<n;j++){
		const int32_t *h = hdr + 8 * (size_t)j;
		const int mlen = h[0], bw = h[1];
		const size_t sz1 = ((size_t)mlen + bw + 15) / 16 * 16, szm = ((size_t)(2 * mlen + 1) * (bw + 2) + 15) / 16 * 16;
		const uint8_t *a = in + in_off[j] + bw / 2;
		uint8_t *M0 = calloc(szm + 64, 1), *M1 = calloc(szm + 64, 1);
		int32_t scr = 0;
		if(getenv("REMSA_DEBUG")) fprintf(stderr, "job %u: mlen %d bw %d mbeg %d mend %d rend %d\n", j, mlen, bw, h[2], h[3], h[4]);
		int err = bso_remsa_core(mlen, bw, h[2], h[3], h[4], a, a + sz1, a + 2 * sz1, a + 3 * sz1, a + 4 * sz1, a + 5 * sz1, a + 6 * sz1, a + 7 * sz1, a + 8 * sz1, a + 9 * sz1,
			M0, M1, match + match_off[j], &scr);
		if(getenv("REMSA_DEBUG")) fprintf(stderr, "job %u: scr %d err %d\n", j, scr, err);
		out[4 * j] = scr; out[4 * j + 1] = err; out[4 * j + 2] = 0; out[4 * j + 3] = 0;
		free(M0); free(M1);
	}
	return 0;
}
#define B200_REMSA_BATCH_FN cpu_remsa_batch
#endif
#include "bspoa.h"
#include "bsalign_b200_poa_compat.h"
#include <time.h>

static uint64_t rng_state;
static inline uint64_t rng_next(void){ rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27; return rng_state * 2685821657736338717ULL; }
static inline double rng_unif(void){ return (rng_next() >> 11) * (1.0 / 9007199254740992.0); }
static double now_s(void){ struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

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

static volatile u4i ref_next; static u4i ref_n; static BSPOA **ref_jobs;
static void *ref_worker(void *arg){ (void)arg; while(1){ u4i j = __sync_fetch_and_add(&ref_next, 1); if(j >= ref_n) break; end_bspoa(ref_jobs[j]); } return NULL; }
static void ref_all(BSPOA **gs, u4i n, int nthr){
	pthread_t th[256]; int t;
	ref_next = 0; ref_n = n; ref_jobs = gs;
	if(nthr > 256) nthr = 256;
	for(t=1;t<nthr;t++) pthread_create(&th[t], NULL, ref_worker, NULL);
	ref_worker(NULL);
	for(t=1;t<nthr;t++) pthread_join(th[t], NULL);
}

int main(int argc, char **argv){
	u4i njobs = argc > 1 ? atoi(argv[1]) : 4, nreads = argc > 2 ? atoi(argv[2]) : 12, tlen = argc > 3 ? atoi(argv[3]) : 1500, j, r, bad = 0;
	u4i vary = argc > 6 ? atoi(argv[6]) : 0;
	BSPOAPar par = DEFAULT_BSPOA_PAR;
	BSPOA **ga, **gb;
	u1i *tmpl, *buf;
	double t0, t_ref, t_b;
	int nthr = (int)sysconf(_SC_NPROCESSORS_ONLN), realn;
	bsb200_ctx *ctx = NULL;
	rng_state = (argc > 4 ? strtoull(argv[4], NULL, 10) : 1) * 0x9E3779B97F4A7C15ULL + 88172645463325252ULL;
	if(argc > 5) par.realn = atoi(argv[5]);
	realn = par.realn;
	if(argc > 7){ b200_poa_remsa_min_objects = atoi(argv[7]); b200_poa_host_threads = 0; }
#ifdef REMSA_CPU_CHECK
	par.shuffle = 0;   /* arm B runs the tail of end_bspoa twice (realn = 0, then the rounds): the read order must not be restored twice */
#endif
#ifndef REMSA_CPU_CHECK
	ctx = bsb200_create(0, 0);
	if(ctx == NULL){ fprintf(stderr, "poa_remsa_dropin: no CUDA device (the product has no CPU fallback)\n"); return 2; }
#endif
	ga = malloc(sizeof(BSPOA*) * njobs); gb = malloc(sizeof(BSPOA*) * njobs);
	tmpl = malloc(tlen); buf = malloc(2 * (size_t)tlen + 16);
	for(j=0;j<njobs;j++){
		u4i nr = nreads + (vary ? (u4i)(rng_next() % (vary + 1)) : 0);
		ga[j] = init_bspoa(par); gb[j] = init_bspoa(par);
		beg_bspoa(ga[j]); beg_bspoa(gb[j]);
		for(r=0;r<tlen;r++) tmpl[r] = rng_next() & 3;
		for(r=0;r<nr;r++){
			u4i len = mutate(tmpl, tlen, buf, 0.03, 0.03, 0.04);
			fwdbitseqpush_bspoa(ga[j], buf, len);
			fwdbitseqpush_bspoa(gb[j], buf, len);
		}
	}
	t0 = now_s();
	ref_all(ga, njobs, nthr);
	t_ref = now_s() - t0;
	if(getenv("REMSA_DEBUG")) fprintf(stderr, "arm A done in %.3f s\n", t_ref);
	t0 = now_s();
#ifdef REMSA_CPU_CHECK
	for(j=0;j<njobs;j++) gb[j]->par->realn = 0;
	ref_all(gb, njobs, nthr);
	for(j=0;j<njobs;j++) gb[j]->par->realn = realn;
	if(getenv("REMSA_DEBUG")) fprintf(stderr, "arm B graphs built at %.3f s\n", now_s() - t0);
	b200_poa_realign_batch(NULL, gb, njobs);
#else
	b200_poa_host_threads = nthr;
	b200_end_bspoa_batch(ctx, gb, njobs);
#endif
	t_b = now_s() - t0;
	for(j=0;j<njobs;j++){
		BSPOA *a = ga[j], *b = gb[j];
		int same = a->cns->size == b->cns->size && memcmp(a->cns->buffer, b->cns->buffer, a->cns->size) == 0
			&& a->qlt->size == b->qlt->size && memcmp(a->qlt->buffer, b->qlt->buffer, a->qlt->size) == 0
			&& a->alt->size == b->alt->size && memcmp(a->alt->buffer, b->alt->buffer, a->alt->size) == 0
			&& a->msaidxs->size == b->msaidxs->size && memcmp(a->msaidxs->buffer, b->msaidxs->buffer, a->msaidxs->size * sizeof(u4i)) == 0
			&& a->msacols->size == b->msacols->size && memcmp(a->msacols->buffer, b->msacols->buffer, a->msacols->size) == 0;
		if(!same){ bad ++; fprintf(stderr, "job %u: consensus / MSA differ (cns %u vs %u, msa %u vs %u)\n", j, (u4i)a->cns->size, (u4i)b->cns->size, (u4i)a->msacols->size, (u4i)b->msacols->size); }
	}
	printf("poa_remsa_dropin: jobs=%u reads=%u(+%u) tlen=%u realn=%d  identical=%u/%u  cns_len[0]=%u  remsa_batches=%lu remsa_jobs=%lu  host_threads=%d  reference_s=%.3f  arm_b_s=%.3f  whole_job_speedup=%.2f\n",
		njobs, nreads, vary, tlen, realn, njobs - bad, njobs, (u4i)ga[0]->cns->size, b200_poa_remsa_batches, b200_poa_remsa_jobs, nthr, t_ref, t_b, t_ref / t_b);
	for(j=0;j<njobs;j++){ free_bspoa(ga[j]); free_bspoa(gb[j]); }
	if(ctx) bsb200_destroy(ctx);
	return bad ? 1 : 0;
}
