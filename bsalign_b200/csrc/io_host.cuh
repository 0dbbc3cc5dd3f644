// io_host.cuh -- ingest / egress next to the hot path: sequence files into the reference's 2-bit BaseBank words, and
// the text the reference's command line prints for an aligned pair.  Host code only (compiled into libbsalign_b200.so).
//
//   readseq_filereader      filereader.h:609   FASTA / FASTQ records (plain or .gz), tag = header up to the first blank, lines concatenated
//   seq2basebank            dna.h:653-671      base_bit_table[c] & 3 into 64-bit words, first base in the top two bits (bits2bit, dna.h:63)
//   seqalign_cigar2alnstr   bsalign.h:531-582  the three alignment rows (query, match line, target)
//   result line             main.c:346-366 (align), main.c:226-228 (edit)
//
// With bsb200_batch_upload_bits the words go to the device as they are: the per-pair bitseq_basebank unpack of main.c:316-321 and the
// 1 byte/base copy disappear from the command-line loop.
#pragma once
#include <zlib.h>
#include <string>
#include <vector>

struct bsb200_seqfile {
	std::vector<uint64_t> bits;      // BaseBank words
	std::vector<uint64_t> off;       // base offset of every record
	std::vector<uint32_t> len;
	std::vector<uint64_t> name_off;  // into names (NUL-terminated strings)
	std::string names;
	uint64_t nbases = 0;
	std::string err;
};

static inline uint8_t bsb200_base_code(unsigned char c){
	// base_bit_table of the reference (dna.h): A/a 0, C/c 1, G/g 2, T/t/U/u 3, everything else 4; seq2basebank keeps (code & 3)
	switch(c){
		case 'A': case 'a': return 0;
		case 'C': case 'c': return 1;
		case 'G': case 'g': return 2;
		case 'T': case 't': case 'U': case 'u': return 3;
		default: return 4;
	}
}

extern "C" bsb200_seqfile *bsb200_seqfile_read(const char *path){
	bsb200_seqfile *sf = new bsb200_seqfile();
	gzFile fp = gzopen(path, "rb");   // (reads plain files too)
	if(!fp){ sf->err = std::string("cannot open ") + path; return sf; }
	gzbuffer(fp, 1 << 20);
	std::vector<char> buf(1 << 20);
	std::string line;
	int state = 0;   // 0: between records, 1: FASTA sequence lines, 2..4: FASTQ sequence / plus / quality line
	uint64_t cur_len = 0;
	auto push_bases = [&](const char *p, size_t n){
		for(size_t i=0;i<n;i++){
			const uint64_t c = bsb200_base_code((unsigned char)p[i]) & 3u;
			if((sf->nbases & 31) == 0) sf->bits.push_back(0);
			sf->bits.back() |= c << (62 - 2 * (sf->nbases & 31));
			sf->nbases++;
		}
		cur_len += n;
	};
	auto close_record = [&](){
		if(state == 0) return;
		if(cur_len == 0){   // the reference skips empty records (main.c:312: if(seq->seq->size == 0) continue)
			sf->names.resize(sf->name_off.back()); sf->name_off.pop_back(); sf->off.pop_back();
		} else sf->len.push_back((uint32_t)cur_len);
		state = 0;
	};
	auto open_record = [&](const std::string &hdr){
		size_t e = 1;
		while(e < hdr.size() && hdr[e] != ' ' && hdr[e] != '\t') e++;
		sf->name_off.push_back(sf->names.size());
		sf->names.append(hdr, 1, e - 1); sf->names.push_back('\0');
		sf->off.push_back(sf->nbases);
		cur_len = 0;
	};
	auto handle_line = [&](std::string &ln){
		while(!ln.empty() && (ln.back() == '\n' || ln.back() == '\r')) ln.pop_back();
		if(state == 2){ push_bases(ln.data(), ln.size()); state = 3; return; }
		if(state == 3){ state = 4; return; }                 // '+' line
		if(state == 4){ close_record(); return; }             // quality line
		if(ln.empty()) return;
		if(ln[0] == '>'){ close_record(); open_record(ln); state = 1; }
		else if(ln[0] == '@' && state == 0){ open_record(ln); state = 2; }
		else if(state == 1) push_bases(ln.data(), ln.size());
	};
	int n;
	while((n = gzread(fp, buf.data(), (unsigned)buf.size())) > 0){
		int b = 0;
		for(int i=0;i<n;i++){
			if(buf[i] == '\n'){
				line.append(buf.data() + b, i - b);
				handle_line(line);
				line.clear();
				b = i + 1;
			}
		}
		line.append(buf.data() + b, n - b);
	}
	if(!line.empty()) handle_line(line);
	if(state == 1) close_record();
	else if(state >= 2){ if(cur_len) sf->len.push_back((uint32_t)cur_len); else { sf->names.resize(sf->name_off.back()); sf->name_off.pop_back(); sf->off.pop_back(); } }
	gzclose(fp);
	sf->bits.push_back(0);   // one spare word: readers may fetch the word behind the last base
	return sf;
}

extern "C" const char *bsb200_seqfile_error(const bsb200_seqfile *sf){ return sf ? sf->err.c_str() : "null"; }
extern "C" uint64_t bsb200_seqfile_nseq(const bsb200_seqfile *sf){ return sf ? sf->len.size() : 0; }
extern "C" uint64_t bsb200_seqfile_nbases(const bsb200_seqfile *sf){ return sf ? sf->nbases : 0; }
extern "C" const uint64_t *bsb200_seqfile_bits(const bsb200_seqfile *sf){ return sf->bits.data(); }
extern "C" const uint64_t *bsb200_seqfile_offsets(const bsb200_seqfile *sf){ return sf->off.data(); }
extern "C" const uint32_t *bsb200_seqfile_lengths(const bsb200_seqfile *sf){ return sf->len.data(); }
extern "C" const char *bsb200_seqfile_name(const bsb200_seqfile *sf, uint64_t i){ return sf->names.data() + sf->name_off[i]; }
extern "C" void bsb200_seqfile_free(bsb200_seqfile *sf){ delete sf; }

static inline uint32_t bsb200_bits_base(const uint64_t *bits, uint64_t off){ return (uint32_t)(bits[off >> 5] >> (((~off) & 31u) << 1)) & 3u; }   // bits2bit, dna.h:63

// seqalign_cigar2alnstr (bsalign.h:531-582) on BaseBank words: rows[0] query, rows[1] target, rows[2] match line; each needs rs->aln + 1 bytes.
// Returns the number of columns written.
extern "C" uint32_t bsb200_cigar2alnstr(const uint64_t *bits, uint64_t qoff, uint64_t toff, const bsb200_result_t *rs, const uint32_t *cigar, uint32_t ncigar,
		char *qrow, char *trow, char *mrow, uint32_t length){
	uint32_t z = 0;
	uint64_t x = qoff + (uint32_t)rs->qb, y = toff + (uint32_t)rs->tb;
	for(uint32_t i=0;i<ncigar&&z<length;i++){
		const uint32_t op = cigar[i] & 0xf;
		uint32_t sz = cigar[i] >> 4;
		if(sz > length - z) sz = length - z;
		switch(op){
			case 0: case 7: case 8:
				for(uint32_t j=0;j<sz;j++){ const uint32_t a = bsb200_bits_base(bits, x++), b = bsb200_bits_base(bits, y++); mrow[z] = a == b ? '|' : '*'; qrow[z] = "ACGT"[a]; trow[z] = "ACGT"[b]; z++; }
				break;
			case 1: case 4:
				for(uint32_t j=0;j<sz;j++){ mrow[z] = '-'; qrow[z] = "ACGT"[bsb200_bits_base(bits, x++)]; trow[z] = '-'; z++; }
				break;
			case 2: case 3:
				for(uint32_t j=0;j<sz;j++){ mrow[z] = '-'; qrow[z] = '-'; trow[z] = "ACGT"[bsb200_bits_base(bits, y++)]; z++; }
				break;
			default: break;
		}
	}
	qrow[z] = 0; trow[z] = 0; mrow[z] = 0;
	return z;
}

// The text of one aligned pair as `bsalign align` / `bsalign edit` print it (main.c:346-347, 226-228 and the three rows behind it; nothing
// is printed for a pair without a match, main.c:324).  Appends to `out`; returns the bytes appended.
static inline size_t bsb200_format_pair(std::string &out, const char *qname, uint32_t qlen, const char *tname, uint32_t tlen, const bsb200_result_t *rs,
		const uint64_t *bits, uint64_t qoff, uint64_t toff, const uint32_t *cigar, uint32_t ncigar, std::vector<char> &rows){
	if(rs->mat == 0) return 0;
	const size_t before = out.size();
	const uint32_t L = (uint32_t)rs->aln;
	if(rows.size() < 3 * ((size_t)L + 1)) rows.resize(3 * ((size_t)L + 1));
	char *qrow = rows.data(), *trow = qrow + L + 1, *mrow = trow + L + 1;
	bsb200_cigar2alnstr(bits, qoff, toff, rs, cigar, ncigar, qrow, trow, mrow, L);
	char hdr[1024];
	int k = snprintf(hdr, sizeof(hdr), "\t%d\t+\t%d\t%d\t", (int)qlen, rs->qb, rs->qe);
	out.append(qname); out.append(hdr, k);
	k = snprintf(hdr, sizeof(hdr), "\t%d\t+\t%d\t%d\t%d\t%.3f\t%d\t%d\t%d\t%d\n", (int)tlen, rs->tb, rs->te, rs->score, 1.0 * rs->mat / rs->aln, rs->mat, rs->mis, rs->ins, rs->del);
	out.append(tname); out.append(hdr, k);
	out.append(qrow); out.push_back('\n'); out.append(mrow); out.push_back('\n'); out.append(trow); out.push_back('\n');
	return out.size() - before;
}

// Whole-file command: every two consecutive records of `path` are a pair (main.c:314); they are aligned in batches of `batch_pairs`
// through bsb200_batch_upload_bits and the text of `bsalign align` (kind 0) / `bsalign edit` (kind 1; kind 2 = `edit -m kmer`, bandwidth = k) goes to `out` in input order.
// Returns the number of pairs, or -1 (message via bsb200_last_error).
extern "C" int64_t bsb200_align_file(bsb200_ctx *ctx, int kind, const char *path, int mode, uint32_t bandwidth, const int8_t matrix[16],
		int8_t go1, int8_t ge1, int8_t go2, int8_t ge2, void *out_, uint64_t batch_pairs){
	FILE *out = (FILE*)out_;
	if(!ctx || !path || !out) return -1;
	bsb200_seqfile *sf = bsb200_seqfile_read(path);
	if(!sf->err.empty()){ ctx->err = sf->err; delete sf; return -1; }
	const uint64_t npairs = sf->len.size() / 2;
	if(batch_pairs == 0) batch_pairs = 1u << 20;
	std::vector<uint64_t> qoff, toff;
	std::vector<uint32_t> qlen, tlen, ncg;
	std::vector<bsb200_result_t> res;
	std::vector<uint32_t> cig;
	std::vector<char> rows;
	std::vector<uint8_t> bases;
	std::string text;
	int64_t rc = (int64_t)npairs;
	for(uint64_t p0=0;p0<npairs&&rc>=0;p0+=batch_pairs){
		const uint64_t m = std::min<uint64_t>(batch_pairs, npairs - p0);
		qoff.resize(m); toff.resize(m); qlen.resize(m); tlen.resize(m); ncg.resize(m); res.resize(m);
		uint64_t cap = 0;
		for(uint64_t k=0;k<m;k++){
			qoff[k] = sf->off[2 * (p0 + k)]; qlen[k] = sf->len[2 * (p0 + k)];
			toff[k] = sf->off[2 * (p0 + k) + 1]; tlen[k] = sf->len[2 * (p0 + k) + 1];
			cap += (uint64_t)qlen[k] + tlen[k] + 2;
		}
		cig.resize(cap);
		uint64_t total = 0;
		if(kind == 2){ // `edit -m kmer -k ksz` (main.c:196): the k-mer size travels in `bandwidth`; one base per byte for this entry point
			const uint64_t lo = qoff[0], hi = toff[m - 1] + tlen[m - 1];
			bases.resize(hi - lo + 1);
			for(uint64_t i=lo;i<hi;i++) bases[i - lo] = (uint8_t)((sf->bits[i >> 5] >> (((~i) & 31) << 1)) & 3);
			for(uint64_t k=0;k<m;k++){ qoff[k] -= lo; toff[k] -= lo; }
			const int krc = bsb200_kmer_edit_batch_dense(ctx, m, bases.data(), qoff.data(), qlen.data(), toff.data(), tlen.data(), bandwidth, res.data(), cig.data(), cap, &total, ncg.data(), nullptr);
			for(uint64_t k=0;k<m;k++){ qoff[k] += lo; toff[k] += lo; }
			if(krc){ rc = -1; break; }
		} else {
			bsb200_batch *b = bsb200_batch_upload_bits(ctx, kind, m, sf->bits.data(), qoff.data(), qlen.data(), toff.data(), tlen.data(), mode, bandwidth, matrix, go1, ge1, go2, ge2, 1);
			if(!b){ rc = -1; break; }
			if(bsb200_batch_run(ctx, b) || bsb200_batch_fetch_dense(ctx, b, res.data(), cig.data(), cap, &total, ncg.data(), nullptr)){ bsb200_batch_free(ctx, b); rc = -1; break; }
			bsb200_batch_free(ctx, b);
		}
		uint64_t cpos = 0;
		text.clear();
		for(uint64_t k=0;k<m;k++){
			bsb200_format_pair(text, bsb200_seqfile_name(sf, 2 * (p0 + k)), qlen[k], bsb200_seqfile_name(sf, 2 * (p0 + k) + 1), tlen[k], &res[k],
				sf->bits.data(), qoff[k], toff[k], cig.data() + cpos, ncg[k], rows);
			cpos += ncg[k];
			if(text.size() > (8u << 20)){ fwrite(text.data(), 1, text.size(), out); text.clear(); }
		}
		fwrite(text.data(), 1, text.size(), out);
	}
	fflush(out);
	delete sf;
	return rc;
}

// the formatter alone, for callers (and tests) that hold results already: appends to a caller buffer, returns bytes needed (nothing written past cap)
extern "C" uint64_t bsb200_format_pair_text(char *out, uint64_t cap, const char *qname, uint32_t qlen, const char *tname, uint32_t tlen, const bsb200_result_t *rs,
		const uint64_t *bits, uint64_t qoff, uint64_t toff, const uint32_t *cigar, uint32_t ncigar){
	std::string s; std::vector<char> rows;
	bsb200_format_pair(s, qname, qlen, tname, tlen, rs, bits, qoff, toff, cigar, ncigar, rows);
	if(out && s.size() <= cap) memcpy(out, s.data(), s.size());
	return s.size();
}

// ---- binary MSA (bspoa.h:1555-1685: dump_binary_msa_bspoa / load_binary_msa_bspoa_core) -------------------------------------------
// Records: 0x81 u32 len, bytes            metadata (optional)
//          0x22 u32 mlen, u32 nseq        then mlen columns of nseq + 1 bytes (the reads' bases 0..3 / 4 = gap, then the consensus base),
//                                          then mlen quality bytes and mlen alternative-base bytes
//          0xFF                           end of one MSA
struct bsb200_msa { uint32_t nseq = 0, mlen = 0; std::vector<uint8_t> cols, qlt, alt; std::string meta; };

extern "C" int bsb200_msa_write(void *out_, uint32_t nseq, uint32_t mlen, const uint8_t *cols, const uint8_t *qlt, const uint8_t *alt, const char *meta, uint32_t metalen){
	FILE *out = (FILE*)out_;
	uint8_t tag;
	if(!out || (mlen && (!cols || !qlt || !alt))) return -1;
	if(meta && metalen){ tag = 0x81; fwrite(&tag, 1, 1, out); fwrite(&metalen, 4, 1, out); fwrite(meta, 1, metalen, out); }
	tag = 0x22; fwrite(&tag, 1, 1, out); fwrite(&mlen, 4, 1, out); fwrite(&nseq, 4, 1, out);
	if(mlen){ fwrite(cols, 1, (size_t)mlen * (nseq + 1), out); fwrite(qlt, 1, mlen, out); fwrite(alt, 1, mlen, out); }
	tag = 0xFF; fwrite(&tag, 1, 1, out);
	return ferror(out) ? -1 : 0;
}

// reads the next MSA of the stream; returns NULL at the end of the file or on a malformed record
extern "C" bsb200_msa *bsb200_msa_read(void *inp_){
	FILE *inp = (FILE*)inp_;
	bsb200_msa *m = new bsb200_msa();
	uint8_t tag; uint32_t a, b; bool any = false;
	while(inp && fread(&tag, 1, 1, inp) == 1){
		if(tag == 0xFF){ if(any) return m; continue; }
		if(tag == 0x81){
			if(fread(&a, 4, 1, inp) != 1) break;
			m->meta.resize(a);
			if(a && fread(&m->meta[0], 1, a, inp) != a) break;
			any = true;
		} else if(tag == 0x22){
			if(fread(&a, 4, 1, inp) != 1 || fread(&b, 4, 1, inp) != 1) break;
			m->mlen = a; m->nseq = b;
			m->cols.resize((size_t)a * (b + 1)); m->qlt.resize(a); m->alt.resize(a);
			if(a && (fread(m->cols.data(), 1, m->cols.size(), inp) != m->cols.size() || fread(m->qlt.data(), 1, a, inp) != a || fread(m->alt.data(), 1, a, inp) != a)) break;
			any = true;
		} else break;
	}
	delete m;
	return nullptr;
}
extern "C" uint32_t bsb200_msa_nseq(const bsb200_msa *m){ return m->nseq; }
extern "C" uint32_t bsb200_msa_mlen(const bsb200_msa *m){ return m->mlen; }
extern "C" const uint8_t *bsb200_msa_cols(const bsb200_msa *m){ return m->cols.data(); }
extern "C" const uint8_t *bsb200_msa_qlt(const bsb200_msa *m){ return m->qlt.data(); }
extern "C" const uint8_t *bsb200_msa_alt(const bsb200_msa *m){ return m->alt.data(); }
extern "C" const char *bsb200_msa_meta(const bsb200_msa *m, uint32_t *len){ if(len) *len = (uint32_t)m->meta.size(); return m->meta.data(); }
extern "C" void bsb200_msa_free(bsb200_msa *m){ delete m; }
