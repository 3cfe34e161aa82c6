// Native input stage: BGZF/BAM streaming reader + the per-region counting loop of `count_snps`
// (restates demuxalot/snp_counter.py:142-276 and the read filter of cellranger_specific.py:13-36 /
// BDRhapsody_specific.py:13-36).  Plain C ABI (include/demux_io.h), no Python, no htslib; zlib only.
//
// Exactness contract: for the built-in read filters the output records are identical, element for element and in
// the same order, to what the Python implementation (demuxalot_b200/counting.py) and the reference produce:
//   * groups keyed by (barcode index, hashed UMI) live in an insertion-ordered table, are closed once the read
//     cursor is SEGMENT_LENGTH past their furthest read end (checked when the 1000-bp segment of the cursor
//     changes, after the current read was added) and are emitted in insertion order;
//   * inside a group, reads with equal (start, end, AS) count once, p_group is the product of the per-read
//     misalignment probabilities in read order, SNP positions are visited in first-seen order and every base
//     accumulates prod 0.1^(0.1 min(q, 40)) in double precision (same libm pow as CPython), the 1000x rule and
//     the single-candidate rule decide whether the position yields a call.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/demux_io.h"

namespace {

thread_local std::string g_error;

struct Fail {
    std::string what;
};

// ---------------------------------------------------------------------------------------------- BGZF stream
class Bgzf {
public:
    explicit Bgzf(const char* path) : file_(fopen(path, "rb")) {
        if (!file_) throw Fail{std::string("cannot open ") + path};
    }
    ~Bgzf() {
        if (file_) fclose(file_);
        if (inflater_ready_) inflateEnd(&inflater_);
    }
    void seek(uint64_t voffset) {
        if (fseeko(file_, (off_t)(voffset >> 16), SEEK_SET) != 0) throw Fail{"seek failed"};
        block_.clear();
        cursor_ = 0;
        if (!load_block()) return;
        cursor_ = (size_t)(voffset & 0xFFFF);
    }
    // false at clean EOF before the first byte; throws on truncation
    bool read(void* dst, size_t n) {
        uint8_t* out = (uint8_t*)dst;
        size_t done = 0;
        while (done < n) {
            if (cursor_ >= block_.size()) {
                if (!load_block()) {
                    if (done == 0) return false;
                    throw Fail{"truncated BAM"};
                }
                continue;
            }
            const size_t take = std::min(n - done, block_.size() - cursor_);
            memcpy(out + done, block_.data() + cursor_, take);
            cursor_ += take;
            done += take;
        }
        return true;
    }
    void skip(size_t n) {
        while (n > 0) {
            if (cursor_ >= block_.size()) {
                if (!load_block()) throw Fail{"truncated BAM"};
                continue;
            }
            const size_t take = std::min(n, block_.size() - cursor_);
            cursor_ += take;
            n -= take;
        }
    }

private:
    bool load_block() {
        uint8_t header[18];
        for (;;) {
            const size_t got = fread(header, 1, 12, file_);
            if (got == 0) return false;
            if (got != 12 || header[0] != 0x1F || header[1] != 0x8B) throw Fail{"not a BGZF stream"};
            const unsigned xlen = header[10] | (header[11] << 8);
            std::vector<uint8_t>& extra = extra_;
            extra.resize(xlen);
            if (fread(extra.data(), 1, xlen, file_) != xlen) throw Fail{"truncated BGZF header"};
            int block_size = -1;
            for (size_t p = 0; p + 4 <= xlen;) {
                const unsigned slen = extra[p + 2] | (extra[p + 3] << 8);
                if (extra[p] == 66 && extra[p + 1] == 67) block_size = (extra[p + 4] | (extra[p + 5] << 8)) + 1;
                p += 4 + slen;
            }
            if (block_size < 0) throw Fail{"gzip member without BGZF block size"};
            const size_t payload = (size_t)block_size - 12 - xlen - 8;
            compressed_.resize(payload + 8);
            if (fread(compressed_.data(), 1, payload + 8, file_) != payload + 8) throw Fail{"truncated BGZF block"};
            const uint32_t isize = compressed_[payload + 4] | (compressed_[payload + 5] << 8) |
                                   (compressed_[payload + 6] << 16) | ((uint32_t)compressed_[payload + 7] << 24);
            block_.resize(isize);
            cursor_ = 0;
            if (isize == 0) continue;  // empty block (EOF marker): try the next one
            // one inflate state for the whole stream, reset per block (raw deflate members)
            if (!inflater_ready_) {
                memset(&inflater_, 0, sizeof(inflater_));
                if (inflateInit2(&inflater_, -15) != Z_OK) throw Fail{"zlib init failed"};
                inflater_ready_ = true;
            } else if (inflateReset(&inflater_) != Z_OK) {
                throw Fail{"zlib reset failed"};
            }
            inflater_.next_in = compressed_.data();
            inflater_.avail_in = (uInt)payload;
            inflater_.next_out = block_.data();
            inflater_.avail_out = (uInt)isize;
            if (inflate(&inflater_, Z_FINISH) != Z_STREAM_END) throw Fail{"zlib inflate failed"};
            return true;
        }
    }
    FILE* file_;
    std::vector<uint8_t> compressed_, block_, extra_;
    size_t cursor_ = 0;
    z_stream inflater_;
    bool inflater_ready_ = false;
};

// ---------------------------------------------------------------------------------------------- BAM records
struct Read {
    int32_t ref_id, pos, end;  // end = pos + reference span of the cigar (pos + 1 when the cigar consumes nothing)
    int32_t l_seq;
    uint8_t mapq;
    uint16_t flag;
    std::vector<uint32_t> cigar;
    // views into the record buffer of next_read(): valid until the next record is read
    const uint8_t* seq = nullptr;   // 4-bit packed
    const uint8_t* qual = nullptr;
    const uint8_t* tags = nullptr;
    size_t n_tags = 0;

    char base(int k) const { return "=ACMGRSVTWYHKDBN"[(seq[k >> 1] >> ((k & 1) ? 0 : 4)) & 0xF]; }

    // integer / string tag lookup; returns false when absent
    bool find_tag(const char* name, int64_t* as_int, std::string* as_str) const {
        size_t off = 0;
        const size_t n = n_tags;
        while (off + 3 <= n) {
            const bool hit = tags[off] == (uint8_t)name[0] && tags[off + 1] == (uint8_t)name[1];
            const char kind = (char)tags[off + 2];
            off += 3;
            size_t size = 0;
            int64_t value = 0;
            switch (kind) {
                case 'A': size = 1; value = tags[off]; break;
                case 'c': size = 1; value = (int8_t)tags[off]; break;
                case 'C': size = 1; value = tags[off]; break;
                case 's': { int16_t v; memcpy(&v, &tags[off], 2); size = 2; value = v; break; }
                case 'S': { uint16_t v; memcpy(&v, &tags[off], 2); size = 2; value = v; break; }
                case 'i': { int32_t v; memcpy(&v, &tags[off], 4); size = 4; value = v; break; }
                case 'I': { uint32_t v; memcpy(&v, &tags[off], 4); size = 4; value = v; break; }
                case 'f': size = 4; break;
                case 'Z':
                case 'H': {
                    const size_t len = strnlen((const char*)&tags[off], n - off);
                    if (hit) {
                        if (!as_str) return false;  // a text tag where a number was asked for (e.g. AS:Z): not usable
                        as_str->assign((const char*)&tags[off], len);
                        return true;
                    }
                    off += len + 1;
                    continue;
                }
                case 'B': {
                    const char sub = (char)tags[off];
                    int32_t count;
                    memcpy(&count, &tags[off + 1], 4);
                    const size_t elem = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                    if (hit) return false;  // array tags are never what the filters ask for
                    off += 5 + elem * (size_t)count;
                    continue;
                }
                default: throw Fail{std::string("unknown BAM tag type ") + kind};
            }
            if (hit) {
                if (!as_int) return false;  // a numeric tag where text was asked for (e.g. CB:i): `as_str` stays untouched
                *as_int = value;
                return kind != 'f';
            }
            off += size;
        }
        return false;
    }
};

bool next_read(Bgzf& in, Read& r, std::vector<uint8_t>& scratch) {
    int32_t block_size;
    if (!in.read(&block_size, 4)) return false;
    if (block_size < 32) throw Fail{"corrupt BAM record"};
    scratch.resize((size_t)block_size);
    if (!in.read(scratch.data(), (size_t)block_size)) throw Fail{"truncated BAM (record body missing)"};
    const uint8_t* p = scratch.data();
    int32_t l_seq;
    uint16_t n_cigar;
    memcpy(&r.ref_id, p, 4);
    memcpy(&r.pos, p + 4, 4);
    const uint8_t l_read_name = p[8];
    r.mapq = p[9];
    memcpy(&n_cigar, p + 12, 2);
    memcpy(&r.flag, p + 14, 2);
    memcpy(&l_seq, p + 16, 4);
    r.l_seq = l_seq;
    size_t off = 32 + l_read_name;
    r.cigar.resize(n_cigar);
    if (n_cigar) memcpy(r.cigar.data(), p + off, 4 * (size_t)n_cigar);
    off += 4 * (size_t)n_cigar;
    r.seq = p + off;
    off += (size_t)(l_seq + 1) / 2;
    r.qual = p + off;
    off += (size_t)l_seq;
    if (off > (size_t)block_size) throw Fail{"corrupt BAM record"};
    r.tags = p + off;
    r.n_tags = (size_t)block_size - off;
    int32_t span = 0;
    for (uint32_t c : r.cigar) {
        const unsigned op = c & 0xF;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += (int32_t)(c >> 4);
    }
    r.end = r.pos + (span ? span : 1);
    return true;
}

// ---------------------------------------------------------------------------------------------- counting
struct Observation {
    char base;
    uint8_t quality;
};

struct GroupRead {
    int32_t start, end;
    int64_t score;
    double p_misaligned;
    std::vector<std::pair<int32_t, Observation>> calls;  // SNP calls of this read, in cigar order
};

struct Group {
    int32_t cb;
    int64_t ub;
    int32_t reach;       // furthest reference end
    int32_t min_start, max_end;
    std::vector<GroupRead> reads;
    bool open = true;
};

struct Counter {
    const int64_t* positions;
    int64_t n_positions;
    std::vector<uint8_t> molecules;  // 12-byte records
    std::vector<uint8_t> calls;      // 13-byte records
    int64_t n_molecules = 0, n_calls = 0;

    bool any_in(int64_t start, int64_t end) const {
        const int64_t* lo = std::lower_bound(positions, positions + n_positions, start);
        return lo != positions + n_positions && *lo < end;
    }

    void calls_of_read(const Read& r, std::vector<std::pair<int32_t, Observation>>& out) const {
        out.clear();
        if (!any_in(r.pos, (int64_t)r.end + 1)) return;
        int64_t in_read = 0, in_ref = r.pos;
        for (uint32_t c : r.cigar) {
            const unsigned op = c & 0xF;
            const int64_t len = c >> 4;
            if (op == 0 || op == 7 || op == 8) {
                const int64_t* lo = std::lower_bound(positions, positions + n_positions, in_ref);
                const int64_t* hi = std::lower_bound(positions, positions + n_positions, in_ref + len);
                for (const int64_t* it = lo; it != hi; ++it) {
                    const int64_t k = in_read + (*it - in_ref);
                    if (k < 0 || k >= r.l_seq) throw Fail{"cigar walks past the read"};
                    out.push_back({(int32_t)*it, Observation{r.base((int)k), r.qual[(size_t)k]}});
                }
                in_ref += len;
                in_read += len;
            } else if (op == 2 || op == 3) {
                in_ref += len;
            } else if (op == 1 || op == 4 || op == 5 || op == 6) {
                in_read += len;
            } else {
                throw Fail{"cigar code unknown"};
            }
        }
    }

    static int base_code(char b) {
        switch (b) {
            case 'A': return 0;
            case 'C': return 1;
            case 'G': return 2;
            case 'T': return 3;
            case 'N': return 4;
            default: throw Fail{std::string("base ") + b + " cannot be encoded (A, C, G, T, N only)"};
        }
    }

    // 0.1 ** (0.1 * q) for q = 0..40, each value produced by the same run-time libm pow call CPython makes
    // (snp_counter.py:204); volatile keeps the compiler from folding the calls at build time with other roundings
    double quality_factor[41];
    Counter() {
        for (int q = 0; q <= 40; ++q) {
            volatile double exponent = 0.1 * (double)q;
            quality_factor[q] = pow(0.1, exponent);
        }
    }

    struct Candidate { char base; double p_wrong; };
    struct Item { int32_t position; uint32_t seq; Observation obs; };          // one observation of a group
    struct Emitted { uint32_t first_seq; int32_t position; Candidate call; };   // one call of a molecule
    std::vector<Item> items_;        // scratch of close_group, reused from group to group
    std::vector<Emitted> emitted_;
    std::vector<const GroupRead*> unique_;

    void append_molecule(const Group& g, double p_group) {
        const int32_t molecule = (int32_t)n_molecules++;
        const int32_t ub32 = (int32_t)g.ub;
        const float p_group32 = (float)p_group;
        const size_t m = molecules.size();
        molecules.resize(m + 12);
        memcpy(&molecules[m], &g.cb, 4);
        memcpy(&molecules[m + 4], &ub32, 4);
        memcpy(&molecules[m + 8], &p_group32, 4);
        size_t c = calls.size();
        calls.resize(c + 13 * emitted_.size());
        for (const Emitted& e : emitted_) {
            const uint8_t code = (uint8_t)base_code(e.call.base);
            const float p32 = (float)e.call.p_wrong;
            memcpy(&calls[c], &molecule, 4);
            memcpy(&calls[c + 4], &e.position, 4);
            calls[c + 8] = code;
            memcpy(&calls[c + 9], &p32, 4);
            c += 13;
        }
        n_calls += (int64_t)emitted_.size();
    }

    void close_group(const Group& g) {
        if (!any_in(g.min_start, (int64_t)g.max_end + 1)) return;
        double p_group = 1.0;
        unique_.clear();
        for (const GroupRead& r : g.reads) {
            bool duplicate = false;
            for (const GroupRead* u : unique_)
                if (u->start == r.start && u->end == r.end && u->score == r.score) { duplicate = true; break; }
            if (duplicate) continue;
            unique_.push_back(&r);
            p_group *= r.p_misaligned;
        }
        emitted_.clear();
        if (unique_.size() == 1) {
            // one read: its positions are strictly ascending, so every position has exactly one observation, one
            // candidate, and yields a call with p_wrong = 1 * 0.1^(0.1 min(q, 40)), in the read's own order
            for (const auto& call : unique_[0]->calls)
                emitted_.push_back({0u, call.first,
                                    Candidate{call.second.base, 1 * quality_factor[std::min<int>(call.second.quality, 40)]}});
        } else {
            // several reads: observations sorted by (position, arrival); a position's rank in first-seen order is the
            // arrival number of its first observation, and the calls are emitted in that order
            items_.clear();
            uint32_t seq = 0;
            for (const GroupRead* r : unique_)
                for (const auto& call : r->calls) items_.push_back({call.first, seq++, call.second});
            std::sort(items_.begin(), items_.end(), [](const Item& a, const Item& b) {
                return a.position != b.position ? a.position < b.position : a.seq < b.seq;
            });
            for (size_t lo = 0; lo < items_.size();) {
                size_t hi = lo;
                Candidate candidates[16];  // in first-seen order of the base; a BAM base has 16 possible codes
                int n_candidates = 0;
                for (; hi < items_.size() && items_[hi].position == items_[lo].position; ++hi) {
                    const Observation& o = items_[hi].obs;
                    const double factor = quality_factor[std::min<int>(o.quality, 40)];
                    bool found = false;
                    for (int k = 0; k < n_candidates; ++k)
                        if (candidates[k].base == o.base) { candidates[k].p_wrong *= factor; found = true; break; }
                    if (!found) candidates[n_candidates++] = {o.base, 1 * factor};
                }
                if (n_candidates > 1) {  // the 1000x rule (snp_counter.py:207-210)
                    double best = candidates[0].p_wrong;
                    for (int k = 0; k < n_candidates; ++k) best = std::min(best, candidates[k].p_wrong);
                    int kept = 0;
                    for (int k = 0; k < n_candidates; ++k)
                        if (candidates[k].p_wrong <= best * 1000) candidates[kept++] = candidates[k];
                    n_candidates = kept;
                }
                if (n_candidates == 1) emitted_.push_back({items_[lo].seq, items_[lo].position, candidates[0]});
                lo = hi;
            }
            std::sort(emitted_.begin(), emitted_.end(),
                      [](const Emitted& a, const Emitted& b) { return a.first_seq < b.first_seq; });
        }
        if (emitted_.empty()) return;
        append_molecule(g, p_group);
    }
};

int64_t hash_umi(const std::string& s) {
    // base-5 polynomial modulo the prime 2147483629, evaluated exactly (demuxalot/utils.py:12-22 uses big ints)
    const uint64_t mod = 2147483629ull;
    uint64_t value = 0;
    for (unsigned char ch : s) value = (value * 5 + ch) % mod;
    return (int64_t)value;
}

// (barcode index, hashed UMI) -> slot of the open group: linear probing over one flat array, deletion by backward
// shift (no tombstones), keys packed into 64 bits (the UMI hash is < 2^31, utils.py:12-22)
class OpenIndex {
public:
    OpenIndex() { resize(1024); }
    static uint64_t pack(int32_t cb, int64_t ub) { return ((uint64_t)(uint32_t)cb << 32) | (uint64_t)(uint32_t)ub; }
    // slot of `key`, or -1
    int64_t find(uint64_t key) const {
        for (size_t i = mix(key) & mask_;; i = (i + 1) & mask_) {
            if (keys_[i] == key) return (int64_t)values_[i];
            if (keys_[i] == EMPTY) return -1;
        }
    }
    void insert(uint64_t key, size_t value) {  // key must be absent
        if ((size_ + 1) * 4 > (mask_ + 1) * 3) resize(2 * (mask_ + 1));
        size_t i = mix(key) & mask_;
        while (keys_[i] != EMPTY) i = (i + 1) & mask_;
        keys_[i] = key;
        values_[i] = value;
        ++size_;
    }
    void erase(uint64_t key) {
        size_t i = mix(key) & mask_;
        while (keys_[i] != key) {
            if (keys_[i] == EMPTY) return;
            i = (i + 1) & mask_;
        }
        --size_;
        for (size_t j = (i + 1) & mask_;; j = (j + 1) & mask_) {  // close the gap so that probe chains stay intact
            if (keys_[j] == EMPTY) break;
            const size_t home = mix(keys_[j]) & mask_;
            if (((j - home) & mask_) >= ((j - i) & mask_)) {
                keys_[i] = keys_[j];
                values_[i] = values_[j];
                i = j;
            }
        }
        keys_[i] = EMPTY;
    }

private:
    static constexpr uint64_t EMPTY = ~0ull;  // cb = -1 never occurs: barcode indices are >= 0
    static size_t mix(uint64_t k) {
        k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
        return (size_t)k;
    }
    void resize(size_t capacity) {
        std::vector<uint64_t> old_keys(capacity, EMPTY);
        std::vector<size_t> old_values(capacity);
        old_keys.swap(keys_);
        old_values.swap(values_);
        mask_ = capacity - 1;
        size_ = 0;
        for (size_t i = 0; i < old_keys.size(); ++i)
            if (old_keys[i] != EMPTY) insert(old_keys[i], old_values[i]);
    }
    std::vector<uint64_t> keys_;
    std::vector<size_t> values_;
    size_t mask_ = 0, size_ = 0;
};

}  // namespace

struct dmxio_result {
    std::vector<uint8_t> molecules, calls;
    int64_t n_molecules = 0, n_calls = 0, n_reads_seen = 0;
};

extern "C" {

const char* dmxio_last_error(void) { return g_error.c_str(); }

dmxio_result* dmxio_count_region(const char* bam_path, int32_t ref_id, uint64_t start_voffset, int64_t start,
                                 int64_t stop, const int64_t* positions, int64_t n_positions,
                                 const char* barcode_keys, const int64_t* barcode_key_offsets,
                                 const int32_t* barcode_indices, int64_t n_barcodes, const char* cb_tag, int32_t use_rg,
                                 const char* umi_tag, const char* nhits_tag, const char* score_tag,
                                 int32_t score_diff_max, int32_t mapq_threshold, double p_misaligned_default) {
    try {
        std::unordered_map<std::string, int32_t> whitelist;
        whitelist.reserve((size_t)n_barcodes * 2);
        for (int64_t k = 0; k < n_barcodes; ++k)
            whitelist.emplace(std::string(barcode_keys + barcode_key_offsets[k],
                                          (size_t)(barcode_key_offsets[k + 1] - barcode_key_offsets[k])),
                              barcode_indices[k]);
        const bool check_nhits = nhits_tag && nhits_tag[0];

        Bgzf in(bam_path);
        in.seek(start_voffset);
        Counter counter;
        counter.positions = positions;
        counter.n_positions = n_positions;

        std::vector<Group> groups;                                   // insertion order
        OpenIndex open_index;
        size_t first_open = 0;
        auto flush = [&](double threshold) {
            for (size_t k = first_open; k < groups.size(); ++k) {
                Group& g = groups[k];
                if (g.open && (double)g.reach < threshold) {
                    counter.close_group(g);
                    g.open = false;
                    open_index.erase(OpenIndex::pack(g.cb, g.ub));
                    std::vector<GroupRead>().swap(g.reads);
                }
            }
            while (first_open < groups.size() && !groups[first_open].open) ++first_open;
            if (first_open == groups.size()) { groups.clear(); first_open = 0; }
        };

        std::unique_ptr<dmxio_result> owner(new dmxio_result());  // a Fail thrown below must not leak it
        dmxio_result* const result = owner.get();
        Read read;
        std::vector<uint8_t> scratch;
        std::vector<std::pair<int32_t, Observation>> calls;
        std::string text, rg;
        bool have_segment = false;
        int64_t previous_segment = 0;
        while (next_read(in, read, scratch)) {
            if (read.ref_id != ref_id) {
                if (read.ref_id > ref_id || read.ref_id < 0) break;
                continue;
            }
            if (stop >= 0 && read.pos >= stop) break;
            if (read.flag & 4) continue;
            if (start >= 0 && (int64_t)read.end <= start) continue;
            ++result->n_reads_seen;
            // ---- read filter (cellranger_specific.py / BDRhapsody_specific.py) ----
            int64_t score = 0, nhits = 0;
            if (!read.find_tag(score_tag, &score, nullptr)) throw Fail{std::string("read without ") + score_tag + " tag"};
            if (score <= (int64_t)read.l_seq - score_diff_max) continue;
            if (check_nhits) {
                if (!read.find_tag(nhits_tag, &nhits, nullptr)) throw Fail{std::string("read without ") + nhits_tag + " tag"};
                if (nhits > 1) continue;
            }
            if (!read.find_tag(umi_tag, nullptr, &text)) continue;
            if ((int)read.mapq < mapq_threshold) continue;
            const int64_t ub = hash_umi(text);
            // ---- barcode ----
            if (!read.find_tag(cb_tag, nullptr, &text)) continue;
            if (use_rg) {
                if (!read.find_tag("RG", nullptr, &rg)) throw Fail{"read without RG tag"};
                text.push_back('\x1f');
                text += rg;
            }
            const auto hit = whitelist.find(text);
            if (hit == whitelist.end()) continue;
            const int32_t cb = hit->second;
            // ---- add to its (barcode, UMI) group ----
            counter.calls_of_read(read, calls);
            GroupRead gr{read.pos, read.end, score, p_misaligned_default, calls};
            const uint64_t key = OpenIndex::pack(cb, ub);
            const int64_t slot = open_index.find(key);
            if (slot < 0) {
                open_index.insert(key, groups.size());
                Group g;
                g.cb = cb; g.ub = ub; g.reach = read.end; g.min_start = read.pos; g.max_end = read.end;
                g.reads.push_back(std::move(gr));
                groups.push_back(std::move(g));
            } else {
                Group& g = groups[(size_t)slot];
                g.reach = std::max(g.reach, read.end);
                g.min_start = std::min(g.min_start, read.pos);
                g.max_end = std::max(g.max_end, read.end);
                g.reads.push_back(std::move(gr));
            }
            const int64_t segment = read.pos / 1000;
            if (!have_segment || segment != previous_segment) {
                flush((double)read.pos - 1000.0);
                previous_segment = segment;
                have_segment = true;
            }
        }
        flush(INFINITY);
        result->molecules.swap(counter.molecules);
        result->calls.swap(counter.calls);
        result->n_molecules = counter.n_molecules;
        result->n_calls = counter.n_calls;
        return owner.release();
    } catch (const Fail& f) {
        g_error = f.what;
        return nullptr;
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}

int dmxio_count_coverage(const char* bam_path, int32_t ref_id, uint64_t start_voffset, int64_t start, int64_t stop,
                         const char* umi_tag, const char* nhits_tag, const char* score_tag, int32_t score_diff_max,
                         int32_t mapq_threshold, int32_t quality_threshold, int32_t* counts /* [4, stop - start] */) {
    try {
        const bool check_nhits = nhits_tag && nhits_tag[0];
        const int64_t width = stop - start;
        Bgzf in(bam_path);
        in.seek(start_voffset);
        Read read;
        std::vector<uint8_t> scratch;
        std::string text;
        while (next_read(in, read, scratch)) {
            if (read.ref_id != ref_id) {
                if (read.ref_id > ref_id || read.ref_id < 0) break;
                continue;
            }
            if (read.pos >= stop) break;
            if (read.flag & 4) continue;
            if ((int64_t)read.end <= start) continue;
            // read filter: the built-in parse_read accepts the read (cellranger_specific.py / BDRhapsody_specific.py)
            int64_t score = 0, nhits = 0;
            if (!read.find_tag(score_tag, &score, nullptr)) throw Fail{std::string("read without ") + score_tag + " tag"};
            if (score <= (int64_t)read.l_seq - score_diff_max) continue;
            if (check_nhits) {
                if (!read.find_tag(nhits_tag, &nhits, nullptr)) throw Fail{std::string("read without ") + nhits_tag + " tag"};
                if (nhits > 1) continue;
            }
            if (!read.find_tag(umi_tag, nullptr, &text)) continue;
            if ((int)read.mapq < mapq_threshold) continue;
            if (read.l_seq == 0) continue;
            // aligned bases only (M, =, X); I / S advance the read, D / N the reference, H / P neither
            int64_t qpos = 0, rpos = read.pos;
            for (uint32_t c : read.cigar) {
                const unsigned op = c & 0xF;
                const int64_t len = c >> 4;
                if (op == 0 || op == 7 || op == 8) {
                    const int64_t lo = std::max<int64_t>(rpos, start), hi = std::min<int64_t>(rpos + len, stop);
                    for (int64_t r = lo; r < hi; ++r) {
                        const int64_t q = qpos + (r - rpos);
                        if (quality_threshold && (int)read.qual[(size_t)q] < quality_threshold) continue;
                        int code;
                        switch (read.base((int)q)) {
                            case 'A': code = 0; break;
                            case 'C': code = 1; break;
                            case 'G': code = 2; break;
                            case 'T': code = 3; break;
                            default: continue;
                        }
                        ++counts[code * width + (r - start)];
                    }
                    qpos += len;
                    rpos += len;
                } else if (op == 1 || op == 4) {
                    qpos += len;
                } else if (op == 2 || op == 3) {
                    rpos += len;
                }
            }
        }
        return 0;
    } catch (const Fail& f) {
        g_error = f.what;
        return -1;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

int64_t dmxio_n_molecules(const dmxio_result* r) { return r->n_molecules; }
int64_t dmxio_n_calls(const dmxio_result* r) { return r->n_calls; }
int64_t dmxio_n_reads_seen(const dmxio_result* r) { return r->n_reads_seen; }

void dmxio_copy(const dmxio_result* r, void* molecules_out, void* calls_out) {
    if (!r->molecules.empty()) memcpy(molecules_out, r->molecules.data(), r->molecules.size());
    if (!r->calls.empty()) memcpy(calls_out, r->calls.data(), r->calls.size());
}

void dmxio_free(dmxio_result* r) { delete r; }

}  // extern "C"
