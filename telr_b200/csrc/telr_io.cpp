// telr_io.cpp — host-side I/O either side of the stage-4 device path (include/telr_io.h): BGZF/BAM/BAI reader with indexed
// window queries, `samtools index`, the read gather that packs straight into the telr_af_batch layout, and a sorted-BAM
// writer.  Replaces pysam.fetch + seqtk + SeqIO.index + csplit (TELR_assembly.py:384-471) and samtools view/sort/index
// (TELR_te.py:507-512).  Formats follow the SAM/BAM specification (BGZF blocks, BAM records, BAI bins + 16 kb linear index).
#include <zlib.h>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/telr_io.h"

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------------------------------------ BGZF reader
struct Bgzf {
    int fd = -1;
    int64_t file_size = 0;
    int64_t blk_addr = -1, next_addr = 0;      // compressed offset of the loaded block / of the block after it
    std::vector<uint8_t> raw, data;            // compressed block, inflated payload
    int blk_len = 0, off = 0;                  // payload length, read position inside it
    int64_t n_inflated = 0;
    z_stream zs;
    bool zs_init = false;

    int open_(const char *path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return TELR_IO_EIO;
        struct stat st;
        if (fstat(fd, &st) != 0) return TELR_IO_EIO;
        file_size = st.st_size;
        raw.resize(1 << 16); data.resize(1 << 16);
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) return TELR_IO_ENOMEM;
        zs_init = true;
        return 0;
    }
    void close_()
    {
        if (fd >= 0) ::close(fd);
        fd = -1;
        if (zs_init) inflateEnd(&zs);
        zs_init = false;
    }
    // loads the block at compressed offset addr; 1 = ok, 0 = end of file, < 0 = error
    int load(int64_t addr)
    {
        if (addr == blk_addr) return 1;
        if (addr >= file_size) return 0;
        uint8_t h[18];
        if (pread(fd, h, 18, addr) != 18) return TELR_IO_EIO;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) return TELR_IO_EFORMAT;
        const int xlen = h[10] | h[11] << 8;
        int bsize = -1;
        if (xlen == 6 && h[12] == 'B' && h[13] == 'C') bsize = h[16] | h[17] << 8;
        else {      // BC is not the first extra subfield: walk them
            std::vector<uint8_t> x((size_t)xlen);
            if (pread(fd, x.data(), xlen, addr + 12) != xlen) return TELR_IO_EIO;
            for (int p = 0; p + 4 <= xlen;) {
                const int sl = x[p + 2] | x[p + 3] << 8;
                if (x[p] == 'B' && x[p + 1] == 'C' && sl == 2 && p + 6 <= xlen) bsize = x[p + 4] | x[p + 5] << 8;
                p += 4 + sl;
            }
        }
        if (bsize < 0) return TELR_IO_EFORMAT;
        const int total = bsize + 1, clen = total - xlen - 20;
        if (clen < 0) return TELR_IO_EFORMAT;
        if ((int)raw.size() < total) raw.resize(total);
        if (pread(fd, raw.data(), total, addr) != total) return TELR_IO_EIO;
        const uint32_t isize = raw[total - 4] | raw[total - 3] << 8 | raw[total - 2] << 16 | (uint32_t)raw[total - 1] << 24;
        if (isize > (1u << 16)) return TELR_IO_EFORMAT;
        if (isize) {
            inflateReset(&zs);
            zs.next_in = raw.data() + 12 + xlen; zs.avail_in = (uInt)clen;
            zs.next_out = data.data(); zs.avail_out = (uInt)data.size();
            if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.total_out != isize) return TELR_IO_EFORMAT;
        }
        ++n_inflated;
        blk_addr = addr; next_addr = addr + total; blk_len = (int)isize; off = 0;
        return 1;
    }
    int seek(uint64_t voff)
    {
        int rc = load((int64_t)(voff >> 16));
        if (rc <= 0) return rc < 0 ? rc : TELR_IO_EIO;
        off = (int)(voff & 0xffff);
        return off <= blk_len ? 0 : TELR_IO_EFORMAT;
    }
    uint64_t tell() const { return off >= blk_len && blk_addr >= 0 ? (uint64_t)next_addr << 16 : (uint64_t)blk_addr << 16 | (uint64_t)off; }
    // reads n bytes; returns n, 0 at a clean end of file, < 0 on error / truncation
    int64_t read(void *dst, int64_t n)
    {
        uint8_t *d = (uint8_t *)dst;
        int64_t got = 0;
        while (got < n) {
            if (off >= blk_len) {
                int rc = load(blk_addr < 0 ? 0 : next_addr);
                if (rc < 0) return rc;
                if (rc == 0) return got == 0 ? 0 : TELR_IO_EIO;
                continue;
            }
            const int64_t k = std::min<int64_t>(n - got, blk_len - off);
            memcpy(d + got, data.data() + off, (size_t)k);
            off += (int)k; got += k;
        }
        return got;
    }
};

struct Chunk { uint64_t beg, end; };
struct RefIndex {
    std::unordered_map<uint32_t, std::vector<Chunk>> bins;
    std::vector<uint64_t> lin;
};

int reg2bin(int64_t beg, int64_t end)
{
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}
void reg2bins(int64_t beg, int64_t end, std::vector<uint32_t> &out)
{
    out.clear();
    if (beg >= end) return;
    if (end > (1LL << 29)) end = 1LL << 29;
    --end;
    out.push_back(0);
    for (int64_t k = 1 + (beg >> 26); k <= 1 + (end >> 26); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 9 + (beg >> 23); k <= 9 + (end >> 23); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 73 + (beg >> 20); k <= 73 + (end >> 20); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 585 + (beg >> 17); k <= 585 + (end >> 17); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 4681 + (beg >> 14); k <= 4681 + (end >> 14); ++k) out.push_back((uint32_t)k);
}

bool read_file(const std::string &path, std::vector<uint8_t> &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)n);
    bool ok = n == 0 || fread(out.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

template <class T> T rd(const uint8_t *p) { T v; memcpy(&v, p, sizeof(T)); return v; }

// reference length consumed by a BAM CIGAR (M D N = X)
int64_t cigar_ref_len(const uint8_t *cig, int n)
{
    int64_t l = 0;
    for (int k = 0; k < n; ++k) {
        const uint32_t c = rd<uint32_t>(cig + 4 * k);
        const int op = c & 0xf;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += c >> 4;
    }
    return l;
}

}  // namespace

struct telr_bam {
    Bgzf z;
    std::vector<std::string> ref_name;
    std::vector<int64_t> ref_len;
    std::unordered_map<std::string, int> tid_of;
    std::vector<RefIndex> idx;
    uint64_t first_rec = 0;
    std::string names;              // result buffer of the last fetch
    std::vector<uint8_t> rec;
};

extern "C" const char *telr_io_strerror(int code)
{
    switch (code) {
    case TELR_IO_OK: return "ok";
    case TELR_IO_EINVAL: return "invalid argument";
    case TELR_IO_ENOMEM: return "out of memory";
    case TELR_IO_EIO: return "I/O error or truncated file";
    case TELR_IO_EFORMAT: return "file format not recognised";
    case TELR_IO_ENOINDEX: return "BAM index (.bai) not found";
    case TELR_IO_ECONTIG: return "unknown reference sequence name";
    case TELR_IO_EMISSING: return "read named in the BAM is absent from the raw reads";
    default: return "unknown error";
    }
}

static int bam_read_header(Bgzf &z, std::vector<std::string> &names, std::vector<int64_t> &lens, std::string *text)
{
    uint8_t h[8];
    if (z.read(h, 8) != 8 || memcmp(h, "BAM\1", 4) != 0) return TELR_IO_EFORMAT;
    const int32_t l_text = rd<int32_t>(h + 4);
    std::string t((size_t)std::max(l_text, 0), '\0');
    if (l_text > 0 && z.read(&t[0], l_text) != l_text) return TELR_IO_EIO;
    if (text) *text = t;
    int32_t n_ref;
    if (z.read(&n_ref, 4) != 4 || n_ref < 0) return TELR_IO_EIO;
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name, l_ref;
        if (z.read(&l_name, 4) != 4 || l_name <= 0 || l_name > (1 << 20)) return TELR_IO_EIO;
        std::string nm((size_t)l_name, '\0');
        if (z.read(&nm[0], l_name) != l_name || z.read(&l_ref, 4) != 4) return TELR_IO_EIO;
        nm.resize(strlen(nm.c_str()));
        names.push_back(nm); lens.push_back(l_ref);
    }
    return 0;
}

static int bai_load(const std::string &path, std::vector<RefIndex> &idx)
{
    std::vector<uint8_t> b;
    if (!read_file(path, b)) return TELR_IO_ENOINDEX;
    if (b.size() < 8 || memcmp(b.data(), "BAI\1", 4) != 0) return TELR_IO_EFORMAT;
    size_t p = 4;
    auto need = [&](size_t n) { return p + n <= b.size(); };
    const int32_t n_ref = rd<int32_t>(&b[p]); p += 4;
    idx.assign((size_t)std::max(n_ref, 0), RefIndex());
    for (int r = 0; r < n_ref; ++r) {
        if (!need(4)) return TELR_IO_EFORMAT;
        const int32_t n_bin = rd<int32_t>(&b[p]); p += 4;
        for (int i = 0; i < n_bin; ++i) {
            if (!need(8)) return TELR_IO_EFORMAT;
            const uint32_t bin = rd<uint32_t>(&b[p]); const int32_t n_chunk = rd<int32_t>(&b[p + 4]); p += 8;
            if (n_chunk < 0 || !need((size_t)n_chunk * 16)) return TELR_IO_EFORMAT;
            if (bin != 37450) {         // 37450: metadata pseudo-bin
                auto &v = idx[r].bins[bin];
                v.resize((size_t)n_chunk);
                for (int c = 0; c < n_chunk; ++c) { v[c].beg = rd<uint64_t>(&b[p + 16 * c]); v[c].end = rd<uint64_t>(&b[p + 16 * c + 8]); }
            }
            p += (size_t)n_chunk * 16;
        }
        if (!need(4)) return TELR_IO_EFORMAT;
        const int32_t n_intv = rd<int32_t>(&b[p]); p += 4;
        if (n_intv < 0 || !need((size_t)n_intv * 8)) return TELR_IO_EFORMAT;
        idx[r].lin.resize((size_t)n_intv);
        for (int i = 0; i < n_intv; ++i) idx[r].lin[i] = rd<uint64_t>(&b[p + 8 * i]);
        p += (size_t)n_intv * 8;
    }
    return 0;
}

extern "C" int telr_bam_open(const char *path, telr_bam **out)
{
    if (!path || !out) return TELR_IO_EINVAL;
    *out = nullptr;
    telr_bam *b = new telr_bam();
    int rc = b->z.open_(path);
    if (rc == 0) rc = bam_read_header(b->z, b->ref_name, b->ref_len, nullptr);
    if (rc == 0) {
        b->first_rec = b->z.tell();
        for (size_t i = 0; i < b->ref_name.size(); ++i) b->tid_of[b->ref_name[i]] = (int)i;
        std::string p1 = std::string(path) + ".bai", p2 = path;
        rc = bai_load(p1, b->idx);
        if (rc == TELR_IO_ENOINDEX && p2.size() > 4 && p2.substr(p2.size() - 4) == ".bam") rc = bai_load(p2.substr(0, p2.size() - 4) + ".bai", b->idx);
    }
    if (rc != 0) { b->z.close_(); delete b; return rc; }
    *out = b;
    return 0;
}
extern "C" void telr_bam_close(telr_bam *b) { if (b) { b->z.close_(); delete b; } }
extern "C" int telr_bam_n_ref(const telr_bam *b) { return b ? (int)b->ref_name.size() : 0; }
extern "C" const char *telr_bam_ref_name(const telr_bam *b, int tid) { return b && tid >= 0 && tid < (int)b->ref_name.size() ? b->ref_name[tid].c_str() : nullptr; }
extern "C" int64_t telr_bam_ref_len(const telr_bam *b, int tid) { return b && tid >= 0 && tid < (int)b->ref_len.size() ? b->ref_len[tid] : -1; }
extern "C" int telr_bam_tid(const telr_bam *b, const char *chrom)
{
    if (!b || !chrom) return -1;
    auto it = b->tid_of.find(chrom);
    return it == b->tid_of.end() ? -1 : it->second;
}
extern "C" int64_t telr_bam_blocks_inflated(const telr_bam *b) { return b ? b->z.n_inflated : 0; }

extern "C" int64_t telr_bam_fetch(telr_bam *b, int tid, int64_t beg, int64_t end, const char **names, int64_t *names_bytes)
{
    if (!b || tid < 0 || tid >= (int)b->ref_name.size()) return TELR_IO_ECONTIG;
    b->names.clear();
    if (names) *names = b->names.c_str();
    if (names_bytes) *names_bytes = 0;
    if (beg < 0) beg = 0;
    if (end <= beg || tid >= (int)b->idx.size()) return 0;
    const RefIndex &ri = b->idx[tid];
    // candidate chunks: every bin that can hold an overlapping record, cut by the linear index
    std::vector<uint32_t> bins;
    reg2bins(beg, end, bins);
    uint64_t min_off = 0;
    if (!ri.lin.empty()) {
        size_t w = (size_t)(beg >> 14);
        if (w >= ri.lin.size()) w = ri.lin.size() - 1;
        min_off = ri.lin[w];
        while (min_off == 0 && w > 0) min_off = ri.lin[--w];
    }
    std::vector<Chunk> ch;
    for (uint32_t bin : bins) {
        auto it = ri.bins.find(bin);
        if (it == ri.bins.end()) continue;
        for (const Chunk &c : it->second) if (c.end > min_off) ch.push_back(c);
    }
    if (ch.empty()) return 0;
    std::sort(ch.begin(), ch.end(), [](const Chunk &a, const Chunk &c) { return a.beg < c.beg; });
    std::vector<Chunk> merged;
    for (const Chunk &c : ch) {
        if (!merged.empty() && c.beg <= merged.back().end) merged.back().end = std::max(merged.back().end, c.end);
        else merged.push_back(c);
    }
    int64_t n = 0;
    for (const Chunk &c : merged) {
        int rc = b->z.seek(std::max(c.beg, min_off));
        if (rc != 0) return rc;
        bool past = false;
        while (b->z.tell() < c.end) {
            int32_t bs;
            int64_t g = b->z.read(&bs, 4);
            if (g == 0) break;
            if (g != 4 || bs < 32) return TELR_IO_EIO;
            if ((int)b->rec.size() < bs) b->rec.resize((size_t)bs);
            if (b->z.read(b->rec.data(), bs) != bs) return TELR_IO_EIO;
            const uint8_t *r = b->rec.data();
            const int32_t rid = rd<int32_t>(r), pos = rd<int32_t>(r + 4);
            const int l_name = r[8], n_cig = rd<uint16_t>(r + 12);
            if (rid != tid) { if (rid > tid || rid < 0) { past = true; break; } continue; }
            if (pos >= end) { past = true; break; }
            if (32 + l_name + 4 * n_cig > bs) return TELR_IO_EFORMAT;
            int64_t span = cigar_ref_len(r + 32 + l_name, n_cig);
            if (span < 1) span = 1;
            if (pos + span > beg) { b->names.append((const char *)r + 32, strnlen((const char *)r + 32, (size_t)l_name)); b->names.push_back('\0'); ++n; }
        }
        if (past) break;
    }
    if (names) *names = b->names.data();
    if (names_bytes) *names_bytes = (int64_t)b->names.size();
    return n;
}

// ------------------------------------------------------------------------------------------------ samtools index
extern "C" int telr_bam_index_build(const char *bam_path, const char *bai_path)
{
    if (!bam_path || !bai_path) return TELR_IO_EINVAL;
    Bgzf z;
    int rc = z.open_(bam_path);
    std::vector<std::string> names; std::vector<int64_t> lens;
    if (rc == 0) rc = bam_read_header(z, names, lens, nullptr);
    if (rc != 0) { z.close_(); return rc; }
    const int n_ref = (int)names.size();
    struct Meta { uint64_t beg = 0, end = 0, n_mapped = 0, n_unmapped = 0; bool any = false; };
    std::vector<std::vector<std::pair<uint32_t, Chunk>>> chunks((size_t)n_ref);     // (bin, chunk) in file order
    std::vector<std::vector<uint64_t>> lin((size_t)n_ref);
    std::vector<Meta> meta((size_t)n_ref);
    uint64_t n_no_coor = 0;
    std::vector<uint8_t> rec;
    int last_tid = -1; int64_t last_pos = -1;
    for (;;) {
        const uint64_t v0 = z.tell();
        int32_t bs;
        int64_t g = z.read(&bs, 4);
        if (g == 0) break;
        if (g != 4 || bs < 32) { rc = TELR_IO_EIO; break; }
        if ((int)rec.size() < bs) rec.resize((size_t)bs);
        if (z.read(rec.data(), bs) != bs) { rc = TELR_IO_EIO; break; }
        const uint64_t v1 = z.tell();
        const uint8_t *r = rec.data();
        const int32_t tid = rd<int32_t>(r), pos = rd<int32_t>(r + 4);
        const int l_name = r[8], n_cig = rd<uint16_t>(r + 12), flag = rd<uint16_t>(r + 14);
        if (tid < 0) { ++n_no_coor; continue; }
        if (tid >= n_ref || tid < last_tid || (tid == last_tid && pos < last_pos)) { rc = TELR_IO_EFORMAT; break; }     // not coordinate-sorted
        last_tid = tid; last_pos = pos;
        int64_t span = cigar_ref_len(r + 32 + l_name, n_cig);
        if (span < 1 || (flag & 4)) span = 1;
        const int64_t e = pos + span;
        const uint32_t bin = (uint32_t)reg2bin(pos, e);
        auto &cv = chunks[tid];
        if (!cv.empty() && cv.back().first == bin && cv.back().second.end == v0) cv.back().second.end = v1;
        else cv.push_back({bin, Chunk{v0, v1}});
        auto &lv = lin[tid];
        const size_t w1 = (size_t)((e - 1) >> 14);
        if (lv.size() <= w1) lv.resize(w1 + 1, 0);
        for (size_t w = (size_t)(pos >> 14); w <= w1; ++w) if (lv[w] == 0) lv[w] = v0;
        Meta &m = meta[tid];
        if (!m.any) { m.beg = v0; m.any = true; }
        m.end = v1;
        if (flag & 4) ++m.n_unmapped; else ++m.n_mapped;
    }
    z.close_();
    if (rc != 0) return rc;
    FILE *f = fopen(bai_path, "wb");
    if (!f) return TELR_IO_EIO;
    auto w32 = [&](int32_t v) { fwrite(&v, 4, 1, f); };
    auto wu32 = [&](uint32_t v) { fwrite(&v, 4, 1, f); };
    auto w64 = [&](uint64_t v) { fwrite(&v, 8, 1, f); };
    fwrite("BAI\1", 1, 4, f);
    w32(n_ref);
    for (int t = 0; t < n_ref; ++t) {
        std::vector<std::pair<uint32_t, Chunk>> cv = chunks[t];
        std::stable_sort(cv.begin(), cv.end(), [](const std::pair<uint32_t, Chunk> &a, const std::pair<uint32_t, Chunk> &c) { return a.first < c.first; });
        int n_bin = 0;
        for (size_t i = 0; i < cv.size(); ++i) if (i == 0 || cv[i].first != cv[i - 1].first) ++n_bin;
        w32(n_bin + (meta[t].any ? 1 : 0));
        for (size_t i = 0; i < cv.size();) {
            size_t j = i;
            while (j < cv.size() && cv[j].first == cv[i].first) ++j;
            wu32(cv[i].first); w32((int32_t)(j - i));
            for (size_t k = i; k < j; ++k) { w64(cv[k].second.beg); w64(cv[k].second.end); }
            i = j;
        }
        if (meta[t].any) { wu32(37450); w32(2); w64(meta[t].beg); w64(meta[t].end); w64(meta[t].n_mapped); w64(meta[t].n_unmapped); }
        auto &lv = lin[t];
        for (size_t w = 1; w < lv.size(); ++w) if (lv[w] == 0) lv[w] = lv[w - 1];
        w32((int32_t)lv.size());
        for (uint64_t v : lv) w64(v);
    }
    w64(n_no_coor);
    const bool ok = !ferror(f);
    fclose(f);
    return ok ? 0 : TELR_IO_EIO;
}

// ------------------------------------------------------------------------------------------------ read gather
namespace {

struct NtTab {
    uint8_t t[256];
    NtTab()
    {
        memset(t, 4, sizeof(t));
        const char *s = "ACGTU";
        const int v[5] = {0, 1, 2, 3, 3};
        for (int i = 0; i < 5; ++i) { t[(int)s[i]] = (uint8_t)v[i]; t[(int)s[i] + 32] = (uint8_t)v[i]; }
    }
};
const NtTab NT;

// ASCII -> 2-bit + N mask at base offset `off` (a multiple of 64) of zero-initialised arrays
void pack_ascii(const char *s, int64_t len, int64_t off, uint32_t *seq2, uint32_t *nmask)
{
    uint32_t *w2 = seq2 + (off >> 4), *wn = nmask + (off >> 5);
    int64_t i = 0;
    for (; i + 16 <= len; i += 16) {          // 16 bases -> one code word + 16 mask bits (an ambiguous base stores code 0)
        uint32_t a = 0, n = 0;
#pragma GCC unroll 16
        for (int k = 0; k < 16; ++k) { const uint32_t c = NT.t[(uint8_t)s[i + k]]; a |= (c & 3u & ((c >> 2) - 1u)) << (2 * k); n |= (c >> 2) << k; }
        w2[i >> 4] = a;
        wn[i >> 5] |= n << (i & 16);
    }
    for (; i < len; ++i) {
        const uint8_t c = NT.t[(uint8_t)s[i]];
        if (c < 4) w2[i >> 4] |= (uint32_t)c << (2 * (i & 15));
        else wn[i >> 5] |= 1u << (i & 31);
    }
}

uint32_t x31_hash(const char *s)
{
    uint32_t h = (uint32_t)(uint8_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)(uint8_t)*s;
    return h;
}

// buffered line reader over zlib (reads plain and gzip files alike)
struct LineReader {
    gzFile g = nullptr;
    std::vector<char> buf;
    size_t beg = 0, end = 0;
    bool eof = false;
    bool open_(const char *path)
    {
        g = gzopen(path, "rb");
        if (!g) return false;
        gzbuffer(g, 1 << 20);
        buf.resize(4 << 20);
        return true;
    }
    void close_() { if (g) gzclose(g); g = nullptr; }
    bool fill()
    {
        if (eof) return false;
        if (beg > 0) { memmove(buf.data(), buf.data() + beg, end - beg); end -= beg; beg = 0; }
        if (end == buf.size()) buf.resize(buf.size() * 2);
        const int n = gzread(g, buf.data() + end, (unsigned)std::min<size_t>(buf.size() - end, 1u << 30));
        if (n <= 0) { eof = true; return false; }
        end += (size_t)n;
        return true;
    }
    // next line without its terminator; false at end of file
    bool next(const char **s, size_t *len)
    {
        for (;;) {
            const char *nl = (const char *)memchr(buf.data() + beg, '\n', end - beg);
            if (nl) {
                *s = buf.data() + beg; *len = (size_t)(nl - *s);
                beg = (size_t)(nl - buf.data()) + 1;
                if (*len && (*s)[*len - 1] == '\r') --*len;
                return true;
            }
            if (!fill()) {
                if (beg < end) { *s = buf.data() + beg; *len = end - beg; beg = end; return true; }
                return false;
            }
        }
    }
};

struct URead {                    // one unique read needed by some locus
    std::string name;
    int32_t len = -1;             // -1 until seen in the raw reads
    std::vector<uint32_t> p2, pn; // packed on its own at offset 0 (64-base padded)
    std::string ascii;            // kept only when the side files are written
};

}  // namespace

extern "C" void telr_gather_free(telr_gather_out *o)
{
    if (!o) return;
    free(o->seq2); free(o->nmask); free(o->read_off); free(o->read_len); free(o->read_hash); free(o->locus_read_begin);
    free(o->contig_off); free(o->contig_len); free(o->live_index); free(o->n_names);
    o->seq2 = o->nmask = nullptr; o->read_off = nullptr; o->read_len = nullptr; o->read_hash = nullptr; o->locus_read_begin = nullptr;
    o->contig_off = nullptr; o->contig_len = nullptr; o->live_index = nullptr; o->n_names = nullptr;
}

extern "C" int telr_gather_run(const telr_gather_in *in, telr_gather_out *out)
{
    if (!in || !out || in->n_loci < 0 || !in->bam_path || !in->raw_reads_path) return TELR_IO_EINVAL;
    memset(out, 0, sizeof(*out));
    const int n_loci = in->n_loci;
    const bool keep_ascii = in->reads_dir != nullptr;
    int n_thr = in->n_threads > 0 ? in->n_threads : (int)std::thread::hardware_concurrency();
    if (n_thr < 1) n_thr = 1;
    if (n_thr > 64) n_thr = 64;
    auto fail = [&](int rc, const std::string &msg) { snprintf(out->err, sizeof(out->err), "%s", msg.c_str()); telr_gather_free(out); return rc; };

    // ---- 1. window queries: unique read names per locus (sorted: deterministic FASTA / batch order) ----
    double t0 = now_s();
    telr_bam *bam = nullptr;
    int rc = telr_bam_open(in->bam_path, &bam);
    if (rc != 0) return fail(rc, std::string(in->bam_path) + ": " + telr_io_strerror(rc));
    std::unordered_map<std::string, int32_t> uid_of;
    std::vector<URead> ureads;
    std::vector<std::vector<int32_t>> locus_uids((size_t)n_loci);
    out->n_names = (int32_t *)calloc((size_t)n_loci + 1, 4);
    for (int l = 0; l < n_loci; ++l) {
        const int tid = telr_bam_tid(bam, in->chrom[l]);
        if (tid < 0) { telr_bam_close(bam); return fail(TELR_IO_ECONTIG, std::string("invalid contig `") + in->chrom[l] + "`"); }
        const char *nm; int64_t nb;
        int64_t n = telr_bam_fetch(bam, tid, in->win_beg[l], in->win_end[l], &nm, &nb);
        if (n < 0) { telr_bam_close(bam); return fail((int)n, std::string(in->bam_path) + ": " + telr_io_strerror((int)n)); }
        std::vector<std::string> v;
        v.reserve((size_t)n);
        for (const char *p = nm; p < nm + nb; p += strlen(p) + 1) v.emplace_back(p);
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        out->n_names[l] = (int32_t)v.size();
        auto &lu = locus_uids[l];
        lu.reserve(v.size());
        for (auto &s : v) {
            auto it = uid_of.find(s);
            if (it == uid_of.end()) {
                it = uid_of.emplace(s, (int32_t)ureads.size()).first;
                ureads.emplace_back();
                ureads.back().name = s;
            }
            lu.push_back(it->second);
        }
    }
    out->bgzf_blocks = telr_bam_blocks_inflated(bam);
    telr_bam_close(bam);
    out->unique_reads = (int64_t)ureads.size();
    out->t_bam_s = now_s() - t0;

    // ---- 2. one streaming pass over the raw reads: keep only what some locus needs; worker threads pack it to 2 bits ----
    t0 = now_s();
    {
        LineReader lr;
        if (!lr.open_(in->raw_reads_path)) return fail(TELR_IO_EIO, std::string(in->raw_reads_path) + ": cannot open");
        std::mutex mu;
        std::condition_variable cv;
        std::deque<int32_t> todo;
        bool done = false;
        auto packer = [&]() {
            for (;;) {
                int32_t uid;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return done || !todo.empty(); });
                    if (todo.empty()) return;
                    uid = todo.front(); todo.pop_front();
                }
                URead &u = ureads[uid];
                const int64_t nb = ((int64_t)u.len + 63) / 64 * 64;
                u.p2.assign((size_t)(nb / 16), 0); u.pn.assign((size_t)(nb / 32), 0);
                pack_ascii(u.ascii.data(), u.len, 0, u.p2.data(), u.pn.data());
                if (!keep_ascii) std::string().swap(u.ascii);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < std::max(1, n_thr - 1); ++t) pool.emplace_back(packer);
        auto finish = [&]() { { std::lock_guard<std::mutex> lk(mu); done = true; } cv.notify_all(); for (auto &t : pool) t.join(); };
        const char *s; size_t len;
        int64_t found = 0;
        std::string id;
        bool have = lr.next(&s, &len);
        while (have && found < (int64_t)ureads.size()) {
            if (len == 0) { have = lr.next(&s, &len); continue; }
            const char kind = s[0];
            if (kind != '>' && kind != '@') { finish(); lr.close_(); return fail(TELR_IO_EFORMAT, std::string(in->raw_reads_path) + ": neither FASTA nor FASTQ"); }
            size_t e = 1;
            while (e < len && s[e] != ' ' && s[e] != '\t') ++e;
            id.assign(s + 1, e - 1);
            auto it = uid_of.find(id);
            URead *u = it != uid_of.end() && ureads[it->second].len < 0 ? &ureads[it->second] : nullptr;
            int64_t slen = 0;
            if (kind == '>') {           // sequence lines until the next header
                while ((have = lr.next(&s, &len)) && !(len && s[0] == '>')) { if (u) u->ascii.append(s, len); slen += (int64_t)len; }
            } else {                      // FASTQ: sequence lines until '+', then as many quality characters
                while ((have = lr.next(&s, &len)) && !(len && s[0] == '+')) { if (u) u->ascii.append(s, len); slen += (int64_t)len; }
                int64_t q = 0;
                while (q < slen && (have = lr.next(&s, &len))) q += (int64_t)len;
                have = lr.next(&s, &len);
            }
            ++out->reads_scanned; out->bases_scanned += slen;
            if (u) {
                u->len = (int32_t)slen;
                { std::lock_guard<std::mutex> lk(mu); todo.push_back(it->second); }
                cv.notify_one();
                ++found;
            }
        }
        finish();
        lr.close_();
        if (found < (int64_t)ureads.size())
            for (auto &u : ureads) if (u.len < 0) return fail(TELR_IO_EMISSING, u.name);
    }
    out->t_reads_s = now_s() - t0;

    // ---- 3. the batch, locus by locus: contig, then its reads (every sequence on a 64-base boundary) ----
    t0 = now_s();
    std::vector<int> live;
    for (int l = 0; l < n_loci; ++l) if (in->contig_seq && in->contig_seq[l] && in->contig_len[l] > 0) live.push_back(l);
    const int n_live = (int)live.size();
    int64_t n_reads = 0;
    for (int l : live) n_reads += (int64_t)locus_uids[l].size();
    if (n_reads > INT32_MAX) return fail(TELR_IO_EINVAL, "more than 2^31 reads in one batch");
    out->n_live = n_live; out->n_reads = (int32_t)n_reads;
    out->read_off = (int64_t *)malloc(((size_t)n_reads + 1) * 8); out->read_len = (int32_t *)malloc(((size_t)n_reads + 1) * 4);
    out->read_hash = (uint32_t *)malloc(((size_t)n_reads + 1) * 4); out->locus_read_begin = (int32_t *)malloc(((size_t)n_live + 1) * 4);
    out->contig_off = (int64_t *)malloc(((size_t)n_live + 1) * 8); out->contig_len = (int32_t *)malloc(((size_t)n_live + 1) * 4);
    out->live_index = (int32_t *)malloc(((size_t)n_live + 1) * 4);
    if (!out->read_off || !out->read_len || !out->read_hash || !out->locus_read_begin || !out->contig_off || !out->contig_len || !out->live_index)
        return fail(TELR_IO_ENOMEM, "host allocation failed");
    std::vector<uint32_t> uhash(ureads.size());
    for (size_t i = 0; i < ureads.size(); ++i) uhash[i] = x31_hash(ureads[i].name.c_str());
    int64_t off = 0; int32_t ri = 0;
    for (int j = 0; j < n_live; ++j) {
        const int l = live[j];
        out->live_index[j] = l; out->locus_read_begin[j] = ri;
        out->contig_off[j] = off; out->contig_len[j] = in->contig_len[l];
        off += ((int64_t)in->contig_len[l] + 63) / 64 * 64;
        for (int32_t uid : locus_uids[l]) {
            out->read_off[ri] = off; out->read_len[ri] = ureads[uid].len; out->read_hash[ri] = uhash[uid];
            off += ((int64_t)ureads[uid].len + 63) / 64 * 64;
            ++ri;
        }
    }
    out->locus_read_begin[n_live] = ri;
    out->n_bases = off;
    out->seq2 = (uint32_t *)calloc((size_t)(off / 16) + 4, 4); out->nmask = (uint32_t *)calloc((size_t)(off / 32) + 4, 4);
    if (!out->seq2 || !out->nmask) return fail(TELR_IO_ENOMEM, "host allocation of the packed batch failed");
    {
        std::atomic<int> next{0};
        auto work = [&]() {
            for (;;) {
                const int j = next.fetch_add(1);
                if (j >= n_live) break;
                const int l = live[j];
                pack_ascii(in->contig_seq[l], in->contig_len[l], out->contig_off[j], out->seq2, out->nmask);
                int32_t r = out->locus_read_begin[j];
                for (int32_t uid : locus_uids[l]) {
                    const URead &u = ureads[uid];
                    memcpy(out->seq2 + (out->read_off[r] >> 4), u.p2.data(), u.p2.size() * 4);
                    memcpy(out->nmask + (out->read_off[r] >> 5), u.pn.data(), u.pn.size() * 4);
                    ++r;
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < n_thr; ++t) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }
    out->t_pack_s = now_s() - t0;

    // ---- 4. side files <reads_dir>/<locus>.reads.fa (TELR_assembly.py:429-456, TELR_te.py:615-617) ----
    t0 = now_s();
    if (keep_ascii) {
        std::atomic<int> next{0}, bad{0};
        auto work = [&]() {
            for (;;) {
                const int l = next.fetch_add(1);
                if (l >= n_loci) break;
                const std::string path = std::string(in->reads_dir) + "/" + in->locus_name[l] + ".reads.fa";
                FILE *f = fopen(path.c_str(), "wb");
                if (!f) { bad = 1; continue; }
                for (int32_t uid : locus_uids[l]) {
                    const URead &u = ureads[uid];
                    fputc('>', f); fwrite(u.name.data(), 1, u.name.size(), f); fputc('\n', f);
                    fwrite(u.ascii.data(), 1, u.ascii.size(), f); fputc('\n', f);
                }
                if (ferror(f)) bad = 1;
                fclose(f);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < n_thr; ++t) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        if (bad) return fail(TELR_IO_EIO, std::string(in->reads_dir) + ": cannot write the read files");
    }
    out->t_write_s = now_s() - t0;
    return 0;
}

// ------------------------------------------------------------------------------------------------ BAM writer
namespace {

struct BgzfWriter {
    FILE *f = nullptr;
    std::vector<uint8_t> buf, comp;
    int level = 6;
    bool ok = true;
    bool open_(const char *path, int lvl)
    {
        f = fopen(path, "wb");
        level = lvl < 0 || lvl > 9 ? 6 : lvl;
        buf.reserve(0xff00); comp.resize(1 << 17);
        return f != nullptr;
    }
    void flush_block()
    {
        if (buf.empty()) return;
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { ok = false; return; }
        zs.next_in = buf.data(); zs.avail_in = (uInt)buf.size();
        zs.next_out = comp.data() + 18; zs.avail_out = (uInt)comp.size() - 26;
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { ok = false; deflateEnd(&zs); return; }
        const uint32_t clen = (uint32_t)zs.total_out;
        deflateEnd(&zs);
        const uint32_t bsize = clen + 25;
        static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0};
        memcpy(comp.data(), hdr, 16);
        comp[16] = (uint8_t)(bsize & 0xff); comp[17] = (uint8_t)(bsize >> 8);
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf.data(), (uInt)buf.size()), isz = (uint32_t)buf.size();
        memcpy(comp.data() + 18 + clen, &crc, 4); memcpy(comp.data() + 22 + clen, &isz, 4);
        if (fwrite(comp.data(), 1, clen + 26, f) != clen + 26) ok = false;
        buf.clear();
    }
    void write(const void *p, size_t n)
    {
        const uint8_t *s = (const uint8_t *)p;
        while (n) {
            const size_t k = std::min(n, (size_t)0xff00 - buf.size());
            buf.insert(buf.end(), s, s + k);
            s += k; n -= k;
            if (buf.size() >= 0xff00) flush_block();
        }
    }
    bool close_()
    {
        flush_block();
        static const uint8_t eof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (f) { if (fwrite(eof, 1, 28, f) != 28) ok = false; if (fclose(f) != 0) ok = false; }
        f = nullptr;
        return ok;
    }
};

}  // namespace

extern "C" int telr_bam_write_sorted(const char *path, int32_t n_ref, const char *const *ref_name, const int32_t *ref_len,
                                     const char *header_text_extra, int64_t n_rec, const telr_sam_rec *recs, int32_t level)
{
    if (!path || n_ref < 0 || n_rec < 0 || (n_rec > 0 && !recs)) return TELR_IO_EINVAL;
    // samtools sort: by (tid, pos), unmapped (tid < 0) last; the sort is stable on input order
    std::vector<int64_t> order((size_t)n_rec);
    for (int64_t i = 0; i < n_rec; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        const uint32_t ta = (uint32_t)recs[a].tid, tb = (uint32_t)recs[b].tid;       // -1 -> 0xffffffff sorts last
        if (ta != tb) return ta < tb;
        return recs[a].pos < recs[b].pos;
    });
    BgzfWriter w;
    if (!w.open_(path, level)) return TELR_IO_EIO;
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (int i = 0; i < n_ref; ++i) text += std::string("@SQ\tSN:") + ref_name[i] + "\tLN:" + std::to_string(ref_len[i]) + "\n";
    if (header_text_extra) text += header_text_extra;
    w.write("BAM\1", 4);
    int32_t v = (int32_t)text.size();
    w.write(&v, 4); w.write(text.data(), text.size());
    w.write(&n_ref, 4);
    for (int i = 0; i < n_ref; ++i) {
        v = (int32_t)strlen(ref_name[i]) + 1;
        w.write(&v, 4); w.write(ref_name[i], (size_t)v); w.write(&ref_len[i], 4);
    }
    uint8_t code[256];
    memset(code, 15, sizeof(code));
    { const char *s = "=ACMGRSVTWYHKDBN"; for (int i = 0; i < 16; ++i) { code[(int)s[i]] = (uint8_t)i; if (s[i] >= 'A') code[(int)s[i] + 32] = (uint8_t)i; } }
    std::vector<uint8_t> body;
    for (int64_t oi = 0; oi < n_rec; ++oi) {
        const telr_sam_rec &r = recs[order[oi]];
        body.clear();
        const size_t l_name = strlen(r.qname) + 1;
        if (l_name > 255) { w.close_(); return TELR_IO_EINVAL; }
        int64_t span = cigar_ref_len((const uint8_t *)r.cigar, r.n_cigar);
        if (span < 1) span = 1;
        const int32_t pos = r.tid < 0 ? -1 : r.pos;
        const uint16_t bin = (uint16_t)reg2bin(pos < 0 ? -1 : pos, pos < 0 ? 0 : pos + span);
        int32_t h[8];
        h[0] = r.tid; h[1] = pos;
        h[2] = (int32_t)((uint32_t)bin << 16 | (uint32_t)(r.mapq & 0xff) << 8 | (uint32_t)l_name);
        h[3] = (int32_t)((uint32_t)r.flag << 16 | (uint32_t)(r.n_cigar & 0xffff));
        h[4] = r.l_seq; h[5] = -1; h[6] = -1; h[7] = 0;
        body.insert(body.end(), (uint8_t *)h, (uint8_t *)h + 32);
        body.insert(body.end(), (const uint8_t *)r.qname, (const uint8_t *)r.qname + l_name);
        body.insert(body.end(), (const uint8_t *)r.cigar, (const uint8_t *)r.cigar + 4 * (size_t)r.n_cigar);
        for (int i = 0; i < r.l_seq; i += 2) {
            const uint8_t a = code[(uint8_t)r.seq[i]], b = i + 1 < r.l_seq ? code[(uint8_t)r.seq[i + 1]] : 0;
            body.push_back((uint8_t)(a << 4 | b));
        }
        body.insert(body.end(), (size_t)r.l_seq, (uint8_t)0xff);        // FASTA input: no base qualities ('*')
        if (r.aux && r.l_aux > 0) body.insert(body.end(), r.aux, r.aux + r.l_aux);
        v = (int32_t)body.size();
        w.write(&v, 4); w.write(body.data(), body.size());
    }
    if (!w.close_()) return TELR_IO_EIO;
    return telr_bam_index_build(path, (std::string(path) + ".bai").c_str());
}
