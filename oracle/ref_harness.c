/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C entry points around the UNMODIFIED reference headers, which are compiled from where they
 * lie under /root/reference (see oracle/Makefile: -I$(REF)); no reference source is copied here.
 * The resulting oracle/_ref/libbsref.so is used
 *   (1) to pin oracle/bsalign_oracle.c (our own CPU restatement) and
 *   (2) as the "reference" CPU arm of bench.py (--impl reference, cpu_baseline.kind="reference").
 *
 * Functions wrapped:
 *   banded_striped_epi8_seqalign_pairwise   bsalign.h:3854
 *   striped_seqedit_pairwise                bsalign.h:1046
 * The reference is single-threaded; the *_batch entry points run a pthread pool over independent
 * pairs, one b1v mempool + u4v cigars per thread.
 */
#include "bsalign.h"
#include <pthread.h>
#include <stdint.h>

typedef struct {
	int kind; // 0 = epi8, 1 = edit
	uint64_t n;
	const uint8_t *seqs;
	const uint64_t *qoff, *toff;
	const uint32_t *qlen, *tlen;
	int mode;
	uint32_t bandwidth;
	const int8_t *matrix;
	int8_t go1, ge1, go2, ge2;
	int32_t *results;       // n * 10
	uint32_t *cigars;       // arena, may be NULL
	const uint64_t *cgoff;  // n + 1 offsets into arena (capacity per pair), may be NULL
	uint32_t *ncigar;       // n, may be NULL
	volatile uint64_t next;
	int repeat;
} bsref_job_t;

static void bsref_store_result(int32_t *out, seqalign_result_t *rs){
	out[0] = rs->score; out[1] = rs->qb; out[2] = rs->qe; out[3] = rs->tb; out[4] = rs->te;
	out[5] = rs->mat; out[6] = rs->mis; out[7] = rs->ins; out[8] = rs->del; out[9] = rs->aln;
}

static void* bsref_worker(void *arg){
	bsref_job_t *job = (bsref_job_t*)arg;
	b1v *mempool = adv_init_b1v(1024, 0, WORDSIZE, 0);
	u4v *cigars = init_u4v(64);
	seqalign_result_t rs;
	b1i mtx[16];
	uint64_t i, j, cap;
	int r;
	if(job->matrix) memcpy(mtx, job->matrix, 16);
	while(1){
		i = __sync_fetch_and_add(&job->next, 16);
		if(i >= job->n) break;
		for(j=i;j<i+16&&j<job->n;j++){
			for(r=0;r<job->repeat;r++){
				if(job->kind == 0){
					rs = banded_striped_epi8_seqalign_pairwise((u1i*)job->seqs + job->qoff[j], job->qlen[j], (u1i*)job->seqs + job->toff[j], job->tlen[j],
						mempool, cigars, job->mode, job->bandwidth, mtx, job->go1, job->ge1, job->go2, job->ge2, 0);
				} else {
					rs = striped_seqedit_pairwise((u1i*)job->seqs + job->qoff[j], job->qlen[j], (u1i*)job->seqs + job->toff[j], job->tlen[j],
						job->mode, job->bandwidth, mempool, cigars, 0);
				}
			}
			bsref_store_result(job->results + j * 10, &rs);
			if(job->ncigar) job->ncigar[j] = cigars->size;
			if(job->cigars && job->cgoff){
				cap = job->cgoff[j + 1] - job->cgoff[j];
				if(cap > cigars->size) cap = cigars->size;
				memcpy(job->cigars + job->cgoff[j], cigars->buffer, cap * sizeof(uint32_t));
			}
		}
	}
	free_b1v(mempool);
	free_u4v(cigars);
	return NULL;
}

static int bsref_run(bsref_job_t *job, int nthreads){
	pthread_t *tids;
	int i;
	if(nthreads < 1) nthreads = 1;
	job->next = 0;
	if(job->repeat < 1) job->repeat = 1;
	if(nthreads == 1){
		bsref_worker(job);
		return 0;
	}
	tids = malloc(sizeof(pthread_t) * nthreads);
	for(i=0;i<nthreads;i++) pthread_create(tids + i, NULL, bsref_worker, job);
	for(i=0;i<nthreads;i++) pthread_join(tids[i], NULL);
	free(tids);
	return 0;
}

int bsref_wordsize(void){ return WORDSIZE; }

int bsref_epi8_batch(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat){
	bsref_job_t job;
	memset(&job, 0, sizeof(job));
	job.kind = 0; job.n = n; job.seqs = seqs; job.qoff = qoff; job.qlen = qlen; job.toff = toff; job.tlen = tlen;
	job.mode = mode; job.bandwidth = bandwidth; job.matrix = matrix; job.go1 = go1; job.ge1 = ge1; job.go2 = go2; job.ge2 = ge2;
	job.results = results; job.cigars = cigars; job.cgoff = cgoff; job.ncigar = ncigar; job.repeat = repeat;
	return bsref_run(&job, nthreads);
}

int bsref_edit_batch(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat){
	bsref_job_t job;
	memset(&job, 0, sizeof(job));
	job.kind = 1; job.n = n; job.seqs = seqs; job.qoff = qoff; job.qlen = qlen; job.toff = toff; job.tlen = tlen;
	job.mode = mode; job.bandwidth = bandwidth;
	job.results = results; job.cigars = cigars; job.cgoff = cgoff; job.ncigar = ncigar; job.repeat = repeat;
	return bsref_run(&job, nthreads);
}

/*
 * Debug aid for pinning the forward pass: run one epi8 alignment and copy out, per target row,
 * the band offset (begs), the 17 ubegs ints and the u/e/q bytes in LINEAR band order (de-striped).
 * The carving below mirrors bsalign.h:3875-3913 only to locate the rows inside the caller's pool.
 */
int bsref_epi8_rows(const uint8_t *q, uint32_t qlen, const uint8_t *t, uint32_t tlen, int mode, uint32_t bandwidth,
		const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		int32_t *result, int32_t *begs_out, int32_t *ubegs_out, int8_t *u_out, int8_t *e_out, int8_t *q_out){
	b1v *mempool = adv_init_b1v(1024, 0, WORDSIZE, 0);
	u4v *cigars = init_u4v(64);
	seqalign_result_t rs;
	b1i mtx[16], *memp, *ups, *eps, *qps, *ubs;
	u4i bw, W, i, p;
	int piecewise, *begs;
	memcpy(mtx, matrix, 16);
	rs = banded_striped_epi8_seqalign_pairwise((u1i*)q, qlen, (u1i*)t, tlen, mempool, cigars, mode, bandwidth, mtx, go1, ge1, go2, ge2, 0);
	bsref_store_result(result, &rs);
	bw = bandwidth? bandwidth : qlen;
	bw = roundup_times(bw, WORDSIZE);
	W = bw / WORDSIZE;
	piecewise = banded_striped_epi8_seqalign_get_piecewise(go1, ge1, go2, ge2, bw);
	memp = mempool->buffer + mempool->size + WORDSIZE;
	memp += banded_striped_epi8_seqalign_qprof_size(qlen, bw);
	ups = memp + bw; memp += bw * Int64(tlen + 1);
	eps = NULL; qps = NULL;
	if(piecewise){ eps = memp + bw; memp += bw * Int64(tlen + 1); }
	if(piecewise == 2){ qps = memp + bw; memp += bw * Int64(tlen + 1); }
	ubs = memp + roundup_times((WORDSIZE + 1) * sizeof(int), WORDSIZE); memp += (tlen + 1) * roundup_times((WORDSIZE + 1) * sizeof(int), WORDSIZE);
	memp += roundup_times((WORDSIZE + 1) * sizeof(int), WORDSIZE);
	memp += bw * (piecewise + 1);
	begs = ((int*)memp) + 1;
	for(i=0;i<tlen;i++){
		begs_out[i] = begs[i];
		memcpy(ubegs_out + i * 17, ubs + i * roundup_times((WORDSIZE + 1) * sizeof(int), WORDSIZE), 17 * sizeof(int));
		for(p=0;p<bw;p++){
			u4i idx = (p % W) * WORDSIZE + (p / W);
			u_out[(size_t)i * bw + p] = ups[(size_t)i * bw + idx];
			if(e_out) e_out[(size_t)i * bw + p] = eps? eps[(size_t)i * bw + idx] : 0;
			if(q_out) q_out[(size_t)i * bw + p] = qps? qps[(size_t)i * bw + idx] : 0;
		}
	}
	free_b1v(mempool);
	free_u4v(cigars);
	return piecewise;
}
