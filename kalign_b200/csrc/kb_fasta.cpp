// kb_fasta.cpp -- FASTA in / out at memory speed: the data format on either side of the path
//
//   kb200_fasta_read  <-> read_file_stdin + read_fasta   lib/src/msa_io.c:348,412
//   kb200_fasta_write <-> write_msa_fasta                lib/src/msa_io.c:668
//
// The reference reads with getline + one malloc per line and classifies every character with three
// libc calls; it writes with one fprintf per character.  At C4 size (100 000 sequences in, 100 000 x
// alnlen characters out) that is seconds next to a 0.6 s alignment.  Here the file is mapped, line and
// record boundaries are found by all host threads at once, every record is extracted independently
// into two flat arrays (residues, gap counts), and the output is assembled in one buffer at offsets
// known in advance and written with one call.  Host code only (no GPU work to do); semantics are the
// reference's, character for character:
//   * a line's content ends at its first control character (iscntrl: 0..31, 127 -- this is how '\n',
//     '\r' and anything after a stray tab disappear, msa_io.c:381-386);
//   * a line whose first character is '>' starts a record; the name is the rest of its content;
//   * on the other lines letters (isalpha, "C" locale) are residues, punctuation (ispunct) counts as
//     gap characters in front of the next residue (gaps[len]++), everything else is dropped; every
//     character of those lines is counted in letter_freq[] (msa_io.c:455-470);
//   * output: ">name\n", the row in lines of 60 characters, a final newline when the last line is
//     partial (msa_io.c:679-709).
#include "../../include/kalign_b200.h"

#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

int kb_default_threads();

namespace {

enum : uint8_t { CH_DROP = 0, CH_ALPHA = 1, CH_PUNCT = 2, CH_END = 3 };

struct CharClass {
        uint8_t c[256];
        CharClass()
        {
                for (int i = 0; i < 256; i++) {
                        uint8_t k = CH_DROP;                                   // digits, space, bytes >= 128
                        if (i < 32 || i == 127) {
                                k = CH_END;
                        } else if ((i >= 'A' && i <= 'Z') || (i >= 'a' && i <= 'z')) {
                                k = CH_ALPHA;
                        } else if ((i >= 33 && i <= 47) || (i >= 58 && i <= 64) || (i >= 91 && i <= 96) || (i >= 123 && i <= 126)) {
                                k = CH_PUNCT;
                        }
                        c[i] = k;
                }
        }
};
const CharClass g_cls;

int use_threads(int n_threads)
{
        int t = n_threads > 0 ? n_threads : kb_default_threads();
        return std::max(1, std::min(t, 64));
}

}  // namespace

struct kb200_fasta {
        int n = 0;
        std::vector<char> names;            // NUL-terminated, back to back
        std::vector<int64_t> name_off;
        // residues of record i at seq_off[i] (NUL-terminated), its len + 1 gap counts at the same offset of `gaps`.
        // Slots are sized by the record's byte count in the file (an upper bound of its length), so that one pass
        // over the file can fill them; the slabs are malloc'd and only the used part of a slot is ever touched.
        char* seqs = nullptr;
        int* gaps = nullptr;
        std::vector<int64_t> seq_off;
        std::vector<int> lens;
        std::vector<const char*> seq_ptrs;  // for kb200_kalign
        int letter_freq[128];
        ~kb200_fasta()
        {
                free(seqs);
                free(gaps);
        }
};

namespace {

struct Mapped {
        const unsigned char* p = nullptr;
        size_t size = 0;
        bool mapped = false;
        std::vector<unsigned char> owned;
        ~Mapped()
        {
                if (mapped && p) {
                        munmap((void*)p, size);
                }
        }
};

int load_file(const char* path, Mapped& m)
{
        const int fd = open(path, O_RDONLY);
        if (fd < 0) {
                fprintf(stderr, "[kalign_b200] kb200_fasta_read: cannot open %s: %s\n", path, strerror(errno));
                return KB200_FAIL;
        }
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) {
                m.size = (size_t)st.st_size;
                if (m.size > 0) {
                        void* a = mmap(nullptr, m.size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
                        if (a != MAP_FAILED) {
                                m.p = (const unsigned char*)a;
                                m.mapped = true;
                                madvise(a, m.size, MADV_SEQUENTIAL);
                                close(fd);
                                return KB200_OK;
                        }
                } else {
                        close(fd);
                        return KB200_OK;
                }
        }
        // pipes, or a file system without mmap: read it all
        m.size = 0;
        unsigned char buf[1 << 16];
        for (;;) {
                const ssize_t r = read(fd, buf, sizeof(buf));
                if (r < 0) {
                        if (errno == EINTR) {
                                continue;
                        }
                        close(fd);
                        return KB200_FAIL;
                }
                if (r == 0) {
                        break;
                }
                m.owned.insert(m.owned.end(), buf, buf + r);
        }
        close(fd);
        m.p = m.owned.data();
        m.size = m.owned.size();
        return KB200_OK;
}

// letters: ((ch | 0x20) - 'a') < 26
inline bool all_letters(const unsigned char* q, size_t n)
{
        unsigned bad = 0;
        for (size_t i = 0; i < n; i++) {
                bad |= (unsigned)((unsigned char)((q[i] | 0x20) - 'a') > 25);
        }
        return bad == 0;
}

// body of one record: [b, e) = the bytes after the header line up to the next header (or the end).
// Residues go to seq[], gap counts to gaps[] (gaps[k] is cleared when residue k - 1 is stored, so only
// len + 1 entries are touched), every character of the lines' contents is counted in freq[4][128]
// (four tables: consecutive equal characters would otherwise serialise on one counter).
inline int scan_body(const unsigned char* p, size_t b, size_t e, char* seq, int* gaps, unsigned* freq)
{
        int len = 0;
        size_t i = b;
        gaps[0] = 0;
        while (i < e) {
                const void* nlp = memchr(p + i, 10, e - i);
                const size_t le = nlp ? (size_t)((const unsigned char*)nlp - p) : e;     // end of the raw line
                const size_t n = le - i;
                if (n > 0 && all_letters(p + i, n)) {
                        // the common line: residues only
                        memcpy(seq + len, p + i, n);
                        memset(gaps + len + 1, 0, n * sizeof(int));
                        for (size_t k = 0; k < n; k++) {
                                freq[(k & 3) * 128 + p[i + k]]++;
                        }
                        len += (int)n;
                } else {
                        for (size_t k = i; k < le; k++) {
                                const unsigned char ch = p[k];
                                const uint8_t c = g_cls.c[ch];
                                if (c == CH_END) {
                                        break;                          // the rest of the line is dropped
                                }
                                if (ch < 128) {
                                        freq[ch]++;
                                }
                                if (c == CH_ALPHA) {
                                        seq[len] = (char)ch;
                                        len++;
                                        gaps[len] = 0;
                                } else if (c == CH_PUNCT) {
                                        gaps[len]++;
                                }
                        }
                }
                i = le + 1;
        }
        seq[len] = 0;
        return len;
}

}  // namespace

int kb200_fasta_read(const char* path, int n_threads, kb200_fasta** out)
{
        if (!path || !out) {
                return KB200_FAIL;
        }
        *out = nullptr;
        Mapped m;
        if (load_file(path, m) != KB200_OK) {
                return KB200_FAIL;
        }
        const int T = use_threads(n_threads);
        const unsigned char* p = m.p;
        const size_t size = m.size;
        // 1. header positions: a '>' at offset 0 or right after a '\n'; chunks of the file in parallel
        std::vector<std::vector<size_t>> found((size_t)T);
#pragma omp parallel num_threads(T)
        {
#ifdef _OPENMP
                const int t = omp_get_thread_num();
                const int nt = omp_get_num_threads();
#else
                const int t = 0, nt = 1;
#endif
                const size_t lo = size * (size_t)t / (size_t)nt;
                const size_t hi = size * (size_t)(t + 1) / (size_t)nt;
                std::vector<size_t>& v = found[(size_t)t];
                size_t i = lo;
                while (i < hi) {
                        const void* q = memchr(p + i, '>', hi - i);
                        if (!q) {
                                break;
                        }
                        const size_t pos = (size_t)((const unsigned char*)q - p);
                        if (pos == 0 || p[pos - 1] == '\n') {
                                v.push_back(pos);
                        }
                        i = pos + 1;
                }
        }
        std::vector<size_t> hdr;
        for (const std::vector<size_t>& v : found) {
                hdr.insert(hdr.end(), v.begin(), v.end());
        }
        const size_t n = hdr.size();
        if (n > 0x7fffffffu) {
                fprintf(stderr, "[kalign_b200] kb200_fasta_read: too many records\n");
                return KB200_FAIL;
        }
        // anything that is a residue or a gap character before the first header is an error in the reference
        // ("Encountered a sequence before encountering it's name", msa_io.c:460)
        {
                const size_t first = n ? hdr[0] : size;
                size_t i = 0;
                while (i < first) {
                        for (; i < first; i++) {
                                const uint8_t k = g_cls.c[p[i]];
                                if (k == CH_END) {
                                        break;
                                }
                                if (k == CH_ALPHA || k == CH_PUNCT) {
                                        fprintf(stderr, "[kalign_b200] kb200_fasta_read: %s: sequence data before the first '>' line\n", path);
                                        return KB200_FAIL;
                                }
                        }
                        if (i < first && p[i] != '\n') {
                                const void* nl = memchr(p + i, '\n', first - i);
                                i = nl ? (size_t)((const unsigned char*)nl - p) : first;
                        }
                        i++;
                }
        }
        kb200_fasta* f = new kb200_fasta();
        f->n = (int)n;
        memset(f->letter_freq, 0, sizeof(f->letter_freq));
        f->name_off.resize(n + 1);
        f->seq_off.resize(n + 1);
        f->lens.resize(n);
        f->seq_ptrs.resize(n);
        // 2. slots: names at their exact size (end of the header line's content), residues / gap counts at the
        //    record's byte count (>= its length)
        std::vector<size_t> body(n), name_len(n);
#pragma omp parallel for num_threads(T) schedule(static)
        for (long long s = 0; s < (long long)n; s++) {
                const size_t h = hdr[(size_t)s];
                const size_t e = (size_t)s + 1 < n ? hdr[(size_t)s + 1] : size;
                size_t i = h + 1;
                while (i < e && g_cls.c[p[i]] != CH_END) {
                        i++;
                }
                name_len[(size_t)s] = i - (h + 1);
                if (i < e && p[i] != 10) {
                        const void* nl = memchr(p + i, 10, e - i);
                        i = nl ? (size_t)((const unsigned char*)nl - p) : e;
                }
                body[(size_t)s] = std::min(e, i + 1);
        }
        int64_t no = 0, so = 0;
        for (size_t s = 0; s < n; s++) {
                const size_t e = s + 1 < n ? hdr[s + 1] : size;
                f->name_off[s] = no;
                f->seq_off[s] = so;
                no += (int64_t)name_len[s] + 1;
                so += (int64_t)(e - body[s]) + 1;
        }
        f->name_off[n] = no;
        f->seq_off[n] = so;
        f->names.resize((size_t)no);
        f->seqs = (char*)malloc((size_t)so + 1);
        f->gaps = (int*)malloc(((size_t)so + 1) * sizeof(int));
        if (!f->seqs || !f->gaps) {
                fprintf(stderr, "[kalign_b200] kb200_fasta_read: out of memory\n");
                delete f;
                return KB200_FAIL;
        }
        // 3. one pass over the records, each into its own slots
        std::vector<unsigned> freq((size_t)T * 512, 0);
        int too_long = 0;
#pragma omp parallel for num_threads(T) schedule(dynamic, 64) reduction(| : too_long)
        for (long long s = 0; s < (long long)n; s++) {
#ifdef _OPENMP
                unsigned* fq = freq.data() + (size_t)omp_get_thread_num() * 512;
#else
                unsigned* fq = freq.data();
#endif
                const size_t h = hdr[(size_t)s];
                const size_t e = (size_t)s + 1 < n ? hdr[(size_t)s + 1] : size;
                char* nm = f->names.data() + f->name_off[(size_t)s];
                memcpy(nm, p + h + 1, name_len[(size_t)s]);
                nm[name_len[(size_t)s]] = 0;
                char* sq = f->seqs + f->seq_off[(size_t)s];
                if (e - body[(size_t)s] > 0x7ffffff0u) {
                        too_long = 1;
                        sq[0] = 0;
                        continue;
                }
                f->lens[(size_t)s] = scan_body(p, body[(size_t)s], e, sq, f->gaps + f->seq_off[(size_t)s], fq);
                f->seq_ptrs[(size_t)s] = sq;
        }
        if (too_long) {
                fprintf(stderr, "[kalign_b200] kb200_fasta_read: a record is longer than 2^31 characters\n");
                delete f;
                return KB200_FAIL;
        }
        unsigned long long total_freq[128];
        memset(total_freq, 0, sizeof(total_freq));
        for (size_t t = 0; t < (size_t)T * 4; t++) {
                for (int c = 0; c < 128; c++) {
                        total_freq[c] += freq[t * 128 + (size_t)c];
                }
        }
        for (int c = 0; c < 128; c++) {
                f->letter_freq[c] = (int)total_freq[c];
        }
        *out = f;
        return KB200_OK;
}

int kb200_fasta_numseq(const kb200_fasta* f)
{
        return f ? f->n : -1;
}

int kb200_fasta_get(const kb200_fasta* f, int i, const char** name, const char** seq, int* len, const int** gaps)
{
        if (!f || i < 0 || i >= f->n) {
                return KB200_FAIL;
        }
        if (name) {
                *name = f->names.data() + f->name_off[(size_t)i];
        }
        if (seq) {
                *seq = f->seqs + f->seq_off[(size_t)i];
        }
        if (len) {
                *len = f->lens[(size_t)i];
        }
        if (gaps) {
                *gaps = f->gaps + f->seq_off[(size_t)i];
        }
        return KB200_OK;
}

const int* kb200_fasta_letter_freq(const kb200_fasta* f)
{
        return f ? f->letter_freq : nullptr;
}

int kb200_fasta_arrays(const kb200_fasta* f, const char* const** seqs, const int** lens)
{
        if (!f || !seqs || !lens) {
                return KB200_FAIL;
        }
        *seqs = f->seq_ptrs.data();
        *lens = f->lens.data();
        return KB200_OK;
}

void kb200_fasta_free(kb200_fasta* f)
{
        delete f;
}

int kb200_fasta_write(const char* path, const char* const* names, const char* const* rows, int n, int alnlen, int n_threads)
{
        if (!path || !names || !rows || n < 0 || alnlen < 0) {
                return KB200_FAIL;
        }
        const int T = use_threads(n_threads);
        const size_t row_out = (size_t)alnlen + ((size_t)alnlen + 59) / 60;      // characters + one newline per started line
        std::vector<size_t> off((size_t)n + 1);
        off[0] = 0;
        for (int i = 0; i < n; i++) {
                if (!names[i] || !rows[i]) {
                        return KB200_FAIL;
                }
                off[(size_t)i + 1] = off[(size_t)i] + 1 + strlen(names[i]) + 1 + row_out;
        }
        const size_t total = off[(size_t)n];
        const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) {
                fprintf(stderr, "[kalign_b200] kb200_fasta_write: cannot open %s: %s\n", path, strerror(errno));
                return KB200_FAIL;
        }
        char* buf = total ? (char*)malloc(total) : nullptr;
        if (total && !buf) {
                close(fd);
                return KB200_FAIL;
        }
#pragma omp parallel for num_threads(T) schedule(static)
        for (int i = 0; i < n; i++) {
                char* o = buf + off[(size_t)i];
                const size_t nl = strlen(names[i]);
                *o++ = '>';
                memcpy(o, names[i], nl);
                o += nl;
                *o++ = '\n';
                const char* r = rows[i];
                for (int j = 0; j < alnlen; j += 60) {
                        const int c = std::min(60, alnlen - j);
                        memcpy(o, r + j, (size_t)c);
                        o += c;
                        *o++ = '\n';
                }
        }
        int rc = KB200_OK;
        size_t done = 0;
        while (done < total) {
                const ssize_t w = write(fd, buf + done, std::min(total - done, (size_t)1 << 30));
                if (w < 0) {
                        if (errno == EINTR) {
                                continue;
                        }
                        fprintf(stderr, "[kalign_b200] kb200_fasta_write: %s: %s\n", path, strerror(errno));
                        rc = KB200_FAIL;
                        break;
                }
                done += (size_t)w;
        }
        free(buf);
        if (close(fd) != 0) {
                rc = KB200_FAIL;
        }
        return rc;
}
