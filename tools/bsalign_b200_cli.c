/*
 * bsalign_b200_cli -- `bsalign align` and `bsalign edit` (main.c:253-385, 128-250 of the reference) on top of libbsalign_b200.so:
 * same options for the path that is built (-m -k -W -M -X -O -E -Q -P), same text on stdout.  Host code in C, as the reference's;
 * everything heavy happens behind the C ABI (bsb200_align_file: FASTA/FASTQ(.gz) -> BaseBank words -> GPU batches -> text).
 * No CPU fallback: without a CUDA device it says so and exits 2.
 */
#include "bsalign_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>

static int usage(void){
	fprintf(stderr,
		"usage: bsalign_b200_cli align [-m global|extend|overlap] [-W bandwidth] [-M match] [-X mismatch] [-O gapo] [-E gape] [-Q gapo2] [-P gape2] <pairs.fa[.gz]>\n"
		"       bsalign_b200_cli edit  [-m global|extend|overlap|kmer] [-k kmer_size] [-W bandwidth] <pairs.fa[.gz]>\n"
		"defaults follow the reference: align -m overlap -W 0 -M 2 -X 6 -O 3 -E 2 -Q 0 -P 0; edit -m global -W 0 -k 13\n");
	return 1;
}

int main(int argc, char **argv){
	int kind, c, mode, W = 0, ksz = 13, M = 2, X = 6, O = 3, E = 2, Q = 0, P = 0, q, t;
	int8_t mtx[16];
	bsb200_ctx *ctx;
	int64_t n;
	if(argc < 3) return usage();
	if(strcmp(argv[1], "align") == 0) kind = 0;
	else if(strcmp(argv[1], "edit") == 0) kind = 1;
	else return usage();
	mode = kind == 0 ? 1 : 0;   /* main.c:262 (align: overlap), main.c:142 (edit: global) */
	argc--; argv++;
	while((c = getopt(argc, argv, "m:k:W:M:X:O:E:Q:P:")) != -1){
		switch(c){
			case 'm':
				if(strcasecmp(optarg, "global") == 0) mode = 0;
				else if(strcasecmp(optarg, "overlap") == 0) mode = 1;
				else if(strcasecmp(optarg, "extend") == 0) mode = 2;
				else if(strcasecmp(optarg, "kmer") == 0 && kind == 1) mode = 3;   /* main.c:152 */
				else return usage();
				break;
			case 'W': W = atoi(optarg); break;
			case 'k': ksz = atoi(optarg); break;
			case 'M': M = atoi(optarg); break;
			case 'X': X = atoi(optarg); break;
			case 'O': O = atoi(optarg); break;
			case 'E': E = atoi(optarg); break;
			case 'Q': Q = atoi(optarg); break;
			case 'P': P = atoi(optarg); break;
			default: return usage();
		}
	}
	if(optind >= argc) return usage();
	if(W < 0) W = 0;
	if(kind == 1 && mode == 1 && W){ fprintf(stderr, " ** disable band in bsalign-edit's overlap mode ** \n"); W = 0; }   /* main.c:170-173 */
	for(q=0;q<4;q++) for(t=0;t<4;t++) mtx[q * 4 + t] = (int8_t)(q == t ? M : -X);   /* bsalign.h:323; penalties are negated like main.c:284-288 */
	ctx = bsb200_create(0, 0);
	if(ctx == NULL){ fprintf(stderr, "bsalign_b200_cli: no CUDA device; this build has no CPU fallback\n"); return 2; }
	if(mode == 3){ kind = 2; mode = 0; W = ksz < 1 ? 1 : ksz; }   /* kmer_striped_seqedit_pairwise, main.c:196 */
	n = bsb200_align_file(ctx, kind, argv[optind], mode, (uint32_t)W, mtx, (int8_t)-O, (int8_t)-E, (int8_t)-Q, (int8_t)-P, stdout, 0);
	if(n < 0){ fprintf(stderr, "bsalign_b200_cli: %s\n", bsb200_last_error(ctx)); bsb200_destroy(ctx); return 1; }
	bsb200_destroy(ctx);
	return 0;
}
