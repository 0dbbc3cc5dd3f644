/*
 * oracle/poa_dropin_test.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Drop-in check of include/bsalign_b200_poa_compat.h against the UNMODIFIED reference headers (compiled from where they lie
 * under /root/reference, see oracle/Makefile): the same read sets go through
 *   (a) the reference's own end_bspoa (bspoa.h:4722), and
 *   (b) b200_end_bspoa_batch: all BSPOA objects in lock-step, every read-vs-graph sweep on the GPU (libbsalign_b200.so),
 * and the consensus, its qualities, the alternative bases and the whole MSA matrix must be byte-identical.
 * Usage: poa_dropin <jobs> <reads per job> <template length> <seed> [realn] [host threads, 0 = all cores]
 */
#include "bsalign.h"
#include "bsalign_b200_poa_kmer.h"   /* the band-placement alignments of a round become one GPU batch; the reference arm keeps its own CPU call */
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
static void *ref_worker(void *arg){ while(1){ u4i j = __sync_fetch_and_add(&ref_next, 1); if(j >= ref_n) break; end_bspoa(ref_jobs[j]); } return NULL; }

int main(int argc, char **argv){
	int nthr;
	u4i njobs = argc > 1 ? atoi(argv[1]) : 4, nreads = argc > 2 ? atoi(argv[2]) : 16, tlen = argc > 3 ? atoi(argv[3]) : 2000, j, r, bad = 0;
	BSPOAPar par = DEFAULT_BSPOA_PAR;
	BSPOA **ga, **gb;
	u1i *tmpl, *buf;
	double t0, t_ref, t_gpu;
	bsb200_ctx *ctx;
	rng_state = (argc > 4 ? strtoull(argv[4], NULL, 10) : 1) * 0x9E3779B97F4A7C15ULL + 88172645463325252ULL;
	if(argc > 5) par.realn = atoi(argv[5]);
	ctx = bsb200_create(0, 0);
	if(ctx == NULL){ fprintf(stderr, "poa_dropin: no CUDA device (the product has no CPU fallback)\n"); return 2; }
	ga = malloc(sizeof(BSPOA*) * njobs); gb = malloc(sizeof(BSPOA*) * njobs);
	tmpl = malloc(tlen); buf = malloc(2 * (size_t)tlen + 16);
	for(j=0;j<njobs;j++){
		ga[j] = init_bspoa(par); gb[j] = init_bspoa(par);
		beg_bspoa(ga[j]); beg_bspoa(gb[j]);
		for(r=0;r<tlen;r++) tmpl[r] = rng_next() & 3;
		for(r=0;r<nreads;r++){
			u4i len = mutate(tmpl, tlen, buf, 0.03, 0.03, 0.04);
			fwdbitseqpush_bspoa(ga[j], buf, len);
			fwdbitseqpush_bspoa(gb[j], buf, len);
		}
	}
	/* whole-job arm (6th argument = host threads, 0 = all cores): the reference on that many cores (objects are independent), against the
	   lock-step GPU path with as many host threads sharing the work between the sweeps */
	nthr = argc > 6 ? atoi(argv[6]) : 1;
	if(nthr <= 0) nthr = (int)sysconf(_SC_NPROCESSORS_ONLN);
	b200_poa_host_threads = nthr;
	t0 = now_s();
	if(nthr > 1){
		pthread_t th[256]; int t; ref_next = 0; ref_n = njobs; ref_jobs = ga;
		if(nthr > 256) nthr = 256;
		for(t=1;t<nthr;t++) pthread_create(&th[t], NULL, ref_worker, NULL);
		ref_worker(NULL);
		for(t=1;t<nthr;t++) pthread_join(th[t], NULL);
	} else for(j=0;j<njobs;j++) end_bspoa(ga[j]);
	t_ref = now_s() - t0;
	t0 = now_s();
	b200_end_bspoa_batch(ctx, gb, njobs);
	t_gpu = now_s() - t0;
	for(j=0;j<njobs;j++){
		BSPOA *a = ga[j], *b = gb[j];
		int same = a->cns->size == b->cns->size && memcmp(a->cns->buffer, b->cns->buffer, a->cns->size) == 0
			&& a->qlt->size == b->qlt->size && memcmp(a->qlt->buffer, b->qlt->buffer, a->qlt->size) == 0
			&& a->alt->size == b->alt->size && memcmp(a->alt->buffer, b->alt->buffer, a->alt->size) == 0
			&& a->msaidxs->size == b->msaidxs->size && memcmp(a->msaidxs->buffer, b->msaidxs->buffer, a->msaidxs->size * sizeof(u4i)) == 0
			&& a->msacols->size == b->msacols->size && memcmp(a->msacols->buffer, b->msacols->buffer, a->msacols->size) == 0;
		if(same){   /* the binary MSA: the reference's own writer on its object, the library's writer on ours, byte for byte; and read back */
			char *b1 = NULL, *b2 = NULL; size_t n1 = 0, n2 = 0;
			FILE *f1 = open_memstream(&b1, &n1), *f2 = open_memstream(&b2, &n2);
			dump_binary_msa_bspoa(a, "job", 3, f1); fclose(f1);
			b200_dump_binary_msa_bspoa(b, "job", 3, f2); fclose(f2);
			if(n1 != n2 || memcmp(b1, b2, n1)){ same = 0; fprintf(stderr, "job %u: binary MSA differs (%zu vs %zu bytes)\n", j, n1, n2); }
			else {
				FILE *f3 = fmemopen(b2, n2, "rb");
				bsb200_msa *m = bsb200_msa_read(f3);
				if(!m || bsb200_msa_nseq(m) != a->nrds || bsb200_msa_mlen(m) != a->msaidxs->size){ same = 0; fprintf(stderr, "job %u: binary MSA does not read back\n", j); }
				if(m) bsb200_msa_free(m);
				fclose(f3);
			}
			free(b1); free(b2);
		}
		if(!same){ bad ++; fprintf(stderr, "job %u: consensus / MSA differ (cns %u vs %u, msa %u vs %u)\n", j, (u4i)a->cns->size, (u4i)b->cns->size, (u4i)a->msacols->size, (u4i)b->msacols->size); }
	}
	{
		printf("poa_dropin: jobs=%u reads=%u tlen=%u realn=%d  identical=%u/%u  cns_len[0]=%u msa_bytes[0]=%u  host_threads=%d  reference_s=%.3f  gpu_lockstep_s=%.3f  whole_job_speedup=%.2f  kmer_batches=%lu kmer_pairs=%lu\n",
			njobs, nreads, tlen, par.realn, njobs - bad, njobs, (u4i)ga[0]->cns->size, (u4i)ga[0]->msacols->size, nthr, t_ref, t_gpu, t_ref / t_gpu,
			b200_poa_kmer_batches, b200_poa_kmer_pairs);
	}
	for(j=0;j<njobs;j++){ free_bspoa(ga[j]); free_bspoa(gb[j]); }
	bsb200_destroy(ctx);
	return bad ? 1 : 0;
}
