/*
 * bsalign_b200_poa_kmer.h -- lets b200_end_bspoa_batch (bsalign_b200_poa_compat.h) run the k-mer guided band placement of a whole round
 * (prepare_rd_align_bspoa, bspoa.h:2087-2111: one kmer_striped_seqedit_pairwise of the read against the current consensus per object)
 * as ONE GPU batch (bsb200_kmer_edit_batch) without touching the reference's code.
 *
 * Include it BETWEEN the reference's two headers:
 *     #include "bsalign.h"
 *     #include "bsalign_b200_poa_kmer.h"
 *     #include "bspoa.h"
 *     #include "bsalign_b200_poa_compat.h"
 * From here on the NAME kmer_striped_seqedit_pairwise inside bspoa.h resolves to the hook below.  The hook hands out the result the
 * round runner computed for the calling thread's object (same arguments, same result, same cigar vector); when nothing was prepared -
 * end_bspoa of the reference itself, remsa_pedits_bspoa, any other caller - it calls the reference's own function, so code that does
 * not go through b200_end_bspoa_batch behaves exactly as before.
 */
#ifndef BSALIGN_B200_POA_KMER_H
#define BSALIGN_B200_POA_KMER_H

#include <stdint.h>

typedef struct {
	int armed;
	u4i qlen, tlen;
	seqalign_result_t rs;
	const uint32_t *cigar;
	uint32_t ncigar;
} b200_poa_kmer_slot_t;

static __thread b200_poa_kmer_slot_t *b200_poa_kmer_slot = NULL;
static __thread unsigned long b200_poa_kmer_hits = 0;   /* calls answered from a prepared slot (tests) */

static inline seqalign_result_t b200_poa_kmer_hook(u1i ksz, u1i *qseq, u4i qlen, u1i *tseq, u4i tlen, b1v *mempool, u4v *cigars, int verbose){
	b200_poa_kmer_slot_t *s = b200_poa_kmer_slot;
	if(s && s->armed && s->qlen == qlen && s->tlen == tlen){
		uint32_t i;
		s->armed = 0;
		b200_poa_kmer_hits ++;
		if(cigars){ clear_u4v(cigars); for(i=0;i<s->ncigar;i++) push_u4v(cigars, s->cigar[i]); }
		return s->rs;
	}
	return kmer_striped_seqedit_pairwise(ksz, qseq, qlen, tseq, tlen, mempool, cigars, verbose);
}

#define kmer_striped_seqedit_pairwise b200_poa_kmer_hook

#endif
