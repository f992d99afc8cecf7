// tipsy.cpp — see tipsy.h
#include "tipsy.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

constexpr int kGasFloats = 12, kDarkFloats = 9, kStarFloats = 11;

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

struct RawHeader {
    uint64_t timeBits;
    uint32_t nBodies, nDim, nSph, nDark, nStar;
};

bool readRawHeader(FILE *f, RawHeader &h) {
    unsigned char b[28];
    if (std::fread(b, 1, sizeof(b), f) != sizeof(b)) return false;
    std::memcpy(&h.timeBits, b, 8);
    std::memcpy(&h.nBodies, b + 8, 4);
    std::memcpy(&h.nDim, b + 12, 4);
    std::memcpy(&h.nSph, b + 16, 4);
    std::memcpy(&h.nDark, b + 20, 4);
    std::memcpy(&h.nStar, b + 24, 4);
    return true;
}

int64_t fileSize(FILE *f) {
    if (fseeko(f, 0, SEEK_END) != 0) return -1;
    const int64_t s = (int64_t)ftello(f);
    return s;
}

}  // namespace

void TipsyIO::close() {
    if (file_) std::fclose(file_);
    file_ = nullptr;
}

bool TipsyIO::open(const char *path) {
    close();
    error_.clear();
    file_ = std::fopen(path, "rb");
    if (!file_) return fail(std::string("cannot open ") + path);
    RawHeader h;
    if (!readRawHeader(file_, h)) return fail("file shorter than a tipsy header");
    // host is little-endian (x86-64 / aarch64): a native file has ndim in 1..3 as read, a standard one after a swap
    bigEndian_ = false;
    if (h.nDim < 1 || h.nDim > 3) {
        RawHeader s = h;
        s.timeBits = bswap64(h.timeBits);
        s.nBodies = bswap32(h.nBodies); s.nDim = bswap32(h.nDim);
        s.nSph = bswap32(h.nSph); s.nDark = bswap32(h.nDark); s.nStar = bswap32(h.nStar);
        if (s.nDim < 1 || s.nDim > 3) return fail("not a tipsy file (ndim is neither 1..3 nor its byte swap)");
        h = s;
        bigEndian_ = true;
    }
    if (h.nDim != 3) return fail("tipsy file is not three-dimensional");
    std::memcpy(&time_, &h.timeBits, 8);
    nSph_ = h.nSph; nDark_ = h.nDark; nStar_ = h.nStar;
    nBodies_ = h.nBodies;
    if (nSph_ + nDark_ + nStar_ != nBodies_) return fail("tipsy header: nsph + ndark + nstar != nbodies");
    const int64_t body = 4 * ((int64_t)nSph_ * kGasFloats + (int64_t)nDark_ * kDarkFloats + (int64_t)nStar_ * kStarFloats);
    const int64_t size = fileSize(file_);
    if (size == body + 32) headerBytes_ = 32;
    else if (size == body + 28) headerBytes_ = 28;
    else return fail("tipsy file size does not match its header (truncated or 64-bit positions)");
    return true;
}

bool TipsyIO::load(uint64_t first, uint64_t n, float *x, float *y, float *z) {
    if (!file_) { error_ = "tipsy file not open"; return false; }
    if (first + n > nBodies_) { error_ = "tipsy load: slice past the end of the file"; return false; }
    struct Section { uint64_t begin, count; int floats; int64_t offset; };
    const Section sections[3] = {
        {0, nSph_, kGasFloats, headerBytes_},
        {nSph_, nDark_, kDarkFloats, headerBytes_ + 4 * (int64_t)nSph_ * kGasFloats},
        {nSph_ + nDark_, nStar_, kStarFloats, headerBytes_ + 4 * ((int64_t)nSph_ * kGasFloats + (int64_t)nDark_ * kDarkFloats)},
    };
    constexpr uint64_t kChunk = 1u << 16;   // records per read
    std::vector<uint32_t> buf;
    for (const Section &s : sections) {
        const uint64_t lo = std::max(first, s.begin), hi = std::min(first + n, s.begin + s.count);
        if (lo >= hi) continue;
        if (fseeko(file_, (off_t)(s.offset + 4 * (int64_t)(lo - s.begin) * s.floats), SEEK_SET) != 0) {
            error_ = "tipsy load: seek failed";
            return false;
        }
        buf.resize((size_t)std::min<uint64_t>(kChunk, hi - lo) * s.floats);
        for (uint64_t at = lo; at < hi;) {
            const uint64_t m = std::min<uint64_t>(kChunk, hi - at);
            if (std::fread(buf.data(), 4, (size_t)m * s.floats, file_) != (size_t)m * s.floats) {
                error_ = "tipsy load: short read";
                return false;
            }
            for (uint64_t i = 0; i < m; ++i) {
                const uint32_t *rec = buf.data() + i * s.floats;   // mass, then pos[3]
                uint32_t px = rec[1], py = rec[2], pz = rec[3];
                if (bigEndian_) { px = bswap32(px); py = bswap32(py); pz = bswap32(pz); }
                const uint64_t o = at + i - first;
                std::memcpy(x + o, &px, 4);
                std::memcpy(y + o, &py, 4);
                std::memcpy(z + o, &pz, 4);
            }
            at += m;
        }
    }
    return true;
}

bool TipsyIO::writePositions(const char *path, uint64_t n, const float *x, const float *y, const float *z,
                             bool standard, std::string *error) {
    auto bad = [&](const char *why) { if (error) *error = std::string(why) + " " + path; return false; };
    if (n > 0x7fffffffull) return bad("too many bodies for a tipsy header:");
    FILE *f = std::fopen(path, "wb");
    if (!f) return bad("cannot create");
    unsigned char hdr[32];
    std::memset(hdr, 0, sizeof(hdr));
    uint64_t timeBits = 0;
    uint32_t ints[5] = {(uint32_t)n, 3u, 0u, (uint32_t)n, 0u};
    if (standard) { timeBits = bswap64(timeBits); for (uint32_t &v : ints) v = bswap32(v); }
    std::memcpy(hdr, &timeBits, 8);
    std::memcpy(hdr + 8, ints, 20);
    bool ok = std::fwrite(hdr, 1, 32, f) == 32;
    const float mass = n ? 1.0f / (float)n : 0.0f;
    uint32_t massBits;
    std::memcpy(&massBits, &mass, 4);
    constexpr uint64_t kChunk = 1u << 16;
    std::vector<uint32_t> buf((size_t)std::min<uint64_t>(kChunk, std::max<uint64_t>(n, 1)) * kDarkFloats);
    for (uint64_t at = 0; ok && at < n;) {
        const uint64_t m = std::min<uint64_t>(kChunk, n - at);
        std::fill(buf.begin(), buf.begin() + (size_t)m * kDarkFloats, 0u);
        for (uint64_t i = 0; i < m; ++i) {
            uint32_t *rec = buf.data() + i * kDarkFloats;
            rec[0] = massBits;
            std::memcpy(rec + 1, x + at + i, 4);
            std::memcpy(rec + 2, y + at + i, 4);
            std::memcpy(rec + 3, z + at + i, 4);
            if (standard) for (int k = 0; k < 4; ++k) rec[k] = bswap32(rec[k]);
        }
        ok = std::fwrite(buf.data(), 4, (size_t)m * kDarkFloats, f) == (size_t)m * kDarkFloats;
        at += m;
    }
    ok = (std::fclose(f) == 0) && ok;
    return ok ? true : bad("write failed:");
}
