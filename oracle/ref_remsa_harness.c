/*
 * oracle/ref_remsa_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Captures what the UNMODIFIED reference does inside remsa_pedit_rd_bspoacore (bspoa.h:3916-4045: the anti-diagonal max-match DP
 * maxmat_dp_diag_rowcal, bspoa.h:3856, of one read against the MSA profile, and the walk that merges the read's nodes into MSA columns).
 * That function is called from inside remsa_pedits_bspoa (bspoa.h:4451), so there is no call site of ours to wrap: this translation
 * unit is compiled with -finstrument-functions (oracle/Makefile, target refremsa) and the enter / exit hooks below look at the
 * reference's own buffers whenever the function they are told about is that one:
 *   on entry   the inputs, which live in g->memp in the layout remsa_pedits_bspoa gives them (bspoa.h:4209-4229);
 *   on exit    the two DP matrices (same buffer) and, per read position, the MSA column its node was merged into (the msa nodes of
 *              add_msanodes_bspoa, bspoa.h:3068, share their group header with the read node after merge_nodes_bspoa).
 * The reference headers are compiled from where they lie under /root/reference; no reference source is copied here.
 */
#include "bspoa.h"
#include <stdint.h>

#define NOINST __attribute__((no_instrument_function))

typedef struct { uint8_t *buf; size_t n, cap; } rblob_t;
static rblob_t RB;
static BSPOA *RG = NULL;
static u4i r_next_rid = 0, r_ncall = 0;
static size_t r_hdr_at = 0;

NOINST static void rb_put(const void *p, size_t len){
	size_t pad = (8 - (len & 7)) & 7;
	if(RB.n + len + pad > RB.cap){ RB.cap = (RB.n + len + pad) * 2 + 4096; RB.buf = realloc(RB.buf, RB.cap); }
	memcpy(RB.buf + RB.n, p, len); RB.n += len;
	if(pad){ memset(RB.buf + RB.n, 0, pad); RB.n += pad; }
}

/* the carving of g->memp by remsa_pedits_bspoa, bspoa.h:4209-4229 (only to LOCATE the reference's arrays) */
typedef struct { u4i mlen, bw, HW, sz1, szm; u1i *seqs[2], *matrix[2], *mats[2][4]; } rlay_t;
NOINST static void r_layout(BSPOA *g, rlay_t *L){
	u1i *b = (u1i*)g->memp->buffer;
	int k;
	L->mlen = g->msaidxs->size;
	L->bw = roundup_times(g->par->editbw / 2, WORDSIZE);
	L->HW = L->bw / 2;
	L->sz1 = roundup_times(L->mlen + L->bw, WORDSIZE);
	L->szm = roundup_times((2 * L->mlen + 1) * (L->bw + 2), WORDSIZE);
	L->seqs[0] = b + L->HW; L->seqs[1] = b + L->sz1 + L->HW;
	L->matrix[0] = b + 2 * L->sz1; L->matrix[1] = L->matrix[0] + L->szm;
	for(k=0;k<4;k++){ L->mats[0][k] = b + 2 * L->sz1 + 5 * (size_t)L->szm + L->HW + (size_t)k * L->sz1; L->mats[1][k] = L->mats[0][k] + 4 * (size_t)L->sz1; }
}

/* record: header int32[16] = {magic, rid, mlen, bw, mbeg, mend, rdlen, nevent, ...}, then seqs0, seqs1 (sz1 bytes each, from -HW),
 * mats[0][0..3], mats[1][0..3] (sz1 bytes each, from -HW), then on exit matrix0, matrix1 (szm bytes each), match[rdlen] int32 */
NOINST void __cyg_profile_func_enter(void *fn, void *cs){
	rlay_t L; int32_t hdr[16]; u4i k, mbeg, mend;
	(void)cs;
	if(fn == (void*)remsa_pedits_bspoa){ r_next_rid = 0; if(getenv("REMSA_DEBUG")) fprintf(stderr, "[remsa] round starts\n"); return; }   /* a new round: the reads are visited in order again */
	if(fn != (void*)remsa_pedit_rd_bspoacore || RG == NULL) return;
	r_layout(RG, &L);
	while(r_next_rid < RG->seqs->nseq && RG->seqs->rdlens->buffer[r_next_rid] == 0) r_next_rid ++;
	for(mbeg=0;mbeg<L.mlen&&L.seqs[0][mbeg]>=4;mbeg++);
	for(mend=L.mlen;mend>mbeg&&L.seqs[0][mend-1]>=4;mend--);
	memset(hdr, 0, sizeof(hdr));
	hdr[0] = 0x52454d53; hdr[1] = r_next_rid; hdr[2] = L.mlen; hdr[3] = L.bw; hdr[4] = mbeg; hdr[5] = mend;
	hdr[6] = RG->seqs->rdlens->buffer[r_next_rid]; hdr[8] = L.sz1; hdr[9] = L.szm;
	r_hdr_at = RB.n;
	rb_put(hdr, sizeof(hdr));
	rb_put(L.seqs[0] - L.HW, L.sz1); rb_put(L.seqs[1] - L.HW, L.sz1);
	for(k=0;k<4;k++) rb_put(L.mats[0][k] - L.HW, L.sz1);
	for(k=0;k<4;k++) rb_put(L.mats[1][k] - L.HW, L.sz1);
}

NOINST void __cyg_profile_func_exit(void *fn, void *cs){
	rlay_t L; u4i x, y, b, rid, rdlen, nev = 0; int32_t *match, *hdr;
	(void)cs;
	if(fn != (void*)remsa_pedit_rd_bspoacore || RG == NULL) return;
	r_layout(RG, &L);
	if(getenv("REMSA_DEBUG")) fprintf(stderr, "[remsa] exit, rid counter %u\n", r_next_rid);
	rid = r_next_rid; rdlen = RG->seqs->rdlens->buffer[rid];
	rb_put(L.matrix[0], L.szm); rb_put(L.matrix[1], L.szm);
	match = malloc(sizeof(int32_t) * (rdlen + 1));
	for(x=0;x<rdlen;x++){
		bspoanode_t *v = get_rdnode_bspoa(RG, rid, x);
		match[x] = -1;
		b = v->base;
		if(b < 4){
			for(y=0;y<L.mlen;y++){   /* the msa node of (base, column) that this read node now shares a group with */
				bspoanode_t *u = get_rdnode_bspoa(RG, RG->seqs->nseq + 1 + b, y);
				if(u->header == v->header){ match[x] = y; nev ++; break; }
			}
		}
	}
	rb_put(match, sizeof(int32_t) * rdlen);
	free(match);
	hdr = (int32_t*)(RB.buf + r_hdr_at);
	hdr[7] = nev;
	r_next_rid ++; r_ncall ++;
}

/* one BSPOA job (DEFAULT_BSPOA_PAR with `realn` re-alignment rounds) through the reference's own end_bspoa; every call of
 * remsa_pedit_rd_bspoacore leaves one record in the blob.  Use at most par.seqcore (40) reads: then every read is a row of the MSA. */
NOINST int64_t bsref_remsa_dump(uint32_t nreads, const uint8_t *seqs, const uint64_t *off, const uint32_t *len, int realn, int editbw, uint8_t **out, uint32_t *ncall){
	BSPOAPar par = DEFAULT_BSPOA_PAR;
	BSPOA *g;
	u4i i;
	par.realn = realn;
	if(editbw) par.editbw = editbw;
	g = init_bspoa(par);
	RB.buf = NULL; RB.n = RB.cap = 0; r_ncall = 0;
	beg_bspoa(g);
	for(i=0;i<nreads;i++) fwdbitseqpush_bspoa(g, (u1i*)seqs + off[i], len[i]);
	RG = g; r_next_rid = 0;
	/* end_bspoa runs `realn` rounds; the read counter restarts with every remsa_pedits_bspoa: its first act is add_msanodes_bspoa */
	end_bspoa(g);
	RG = NULL;
	free_bspoa(g);
	*out = RB.buf; if(ncall) *ncall = r_ncall;
	return (int64_t)RB.n;
}
NOINST void bsref_remsa_free(void *p){ free(p); }
