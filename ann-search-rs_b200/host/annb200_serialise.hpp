// annb200_serialise.hpp -- header-only C++17 reader / writer of the crate's saved indices (IndexIo, src/serialise/mod.rs:33-335)
// for the kinds "exhaustive" (src/cpu/exhaustive.rs:18-32) and "ivf" (src/cpu/ivf.rs:24-48), f32.  Same layout and error
// behaviour as python/annb200/serialise.py (see its docstring for the byte layout): the header and the error paths restate
// the reference's own tests (src/serialise/mod.rs:1514-1640); the payload follows bincode 2's "standard" configuration as
// published and is UNPINNED -- no saved index ships with the reference and there is no Rust toolchain in the build image.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/annb200.h"

namespace annb200::serialise {

constexpr char MAGIC[8] = {'A', 'N', 'N', 'S', 'R', 'S', '\0', '\0'};
constexpr uint32_t FORMAT_VERSION = 2;
constexpr const char* INDEX_FILE = "index.bin";

// Serialisation variants of AnnSearchErrors (src/errors.rs): variant() names the enum variant.
class SerialiseError : public std::runtime_error {
  public:
    SerialiseError(std::string variant, const std::string& msg) : std::runtime_error(variant + ": " + msg), variant_(std::move(variant)) {}
    const std::string& variant() const { return variant_; }

  private:
    std::string variant_;
};

struct SavedExhaustive {
    std::vector<float> vectors_flat;   // [n * dim]
    uint64_t dim = 0, n = 0;
    std::vector<float> norms;          // empty for SquaredEuclidean
    uint32_t metric = 0;               // Dist: 0 SquaredEuclidean, 1 Cosine, 2 Manhattan (src/utils/dist.rs:29-37)
};

struct SavedIvf {
    std::vector<float> vectors_flat;   // list order (after optimise_memory_layout, ivf.rs:257-294)
    uint64_t dim = 0, n = 0;
    std::vector<float> norms;
    uint32_t metric = 0;
    std::vector<float> centroids, centroids_norm;
    std::vector<uint64_t> all_indices, offsets;
    uint64_t nlist = 0;
    std::vector<uint64_t> original_ids;
};

namespace detail {

inline void put_varint(std::string& out, uint64_t v) {
    uint8_t buf[9];
    const int64_t n = annb_varint_encode_u64(&v, 1, buf, sizeof(buf));
    out.append(reinterpret_cast<const char*>(buf), static_cast<size_t>(n));
}
inline void put_f32_vec(std::string& out, const std::vector<float>& a) {
    put_varint(out, a.size());
    out.append(reinterpret_cast<const char*>(a.data()), a.size() * 4);   // little-endian hosts only (x86-64 / aarch64)
}
inline void put_usize_vec(std::string& out, const std::vector<uint64_t>& a) {
    put_varint(out, a.size());
    std::vector<uint8_t> buf(a.size() * 9 + 1);
    const int64_t n = annb_varint_encode_u64(a.data(), a.size(), buf.data(), buf.size());
    if (n < 0) throw SerialiseError("EncodeError", "varint encoding failed");
    out.append(reinterpret_cast<const char*>(buf.data()), static_cast<size_t>(n));
}
inline std::string header(const std::string& kind, uint8_t float_width) {
    if (kind.size() > 255)   // one byte holds the tag length (mod.rs:84-92)
        throw SerialiseError("EncodeError", "index kind tag '" + kind + "' is " + std::to_string(kind.size()) + " bytes; the header allows 255");
    std::string h(MAGIC, 8);
    const uint32_t v = FORMAT_VERSION;
    h.append(reinterpret_cast<const char*>(&v), 4);
    h.push_back(static_cast<char>(float_width));
    h.push_back(static_cast<char>(kind.size()));
    h += kind;
    return h;
}

struct Reader {
    const std::string& buf;
    size_t pos;
    void need(size_t n) const {
        if (pos + n > buf.size()) throw SerialiseError("DecodeError", "unexpected end of the payload");
    }
    uint64_t varint() {
        uint64_t v = 0;
        const int64_t used = annb_varint_decode_u64(reinterpret_cast<const uint8_t*>(buf.data()) + pos, buf.size() - pos, 1, &v);
        if (used < 0) throw SerialiseError("DecodeError", "unexpected end of the payload");
        pos += static_cast<size_t>(used);
        return v;
    }
    std::vector<float> f32_vec() {
        const uint64_t n = varint();
        need(n * 4);
        std::vector<float> a(n);
        std::memcpy(a.data(), buf.data() + pos, n * 4);
        pos += n * 4;
        return a;
    }
    std::vector<uint64_t> usize_vec() {
        const uint64_t n = varint();
        if (n > buf.size() - pos) throw SerialiseError("DecodeError", "unexpected end of the payload inside an index list");   // >= 1 byte each
        std::vector<uint64_t> a(n);
        const int64_t used = annb_varint_decode_u64(reinterpret_cast<const uint8_t*>(buf.data()) + pos, buf.size() - pos, n, a.data());
        if (used < 0) throw SerialiseError("DecodeError", "unexpected end of the payload inside an index list");
        pos += static_cast<size_t>(used);
        return a;
    }
    uint32_t dist() {
        const uint64_t v = varint();
        if (v > 2) throw SerialiseError("DecodeError", "unknown Dist variant " + std::to_string(v));
        return static_cast<uint32_t>(v);
    }
};

// read_header (mod.rs:108-170): magic, version, kind, float width, in that order.
inline size_t read_header(const std::string& buf, const std::string& path, const std::string& kind, uint8_t float_width) {
    if (buf.size() < 8 || std::memcmp(buf.data(), MAGIC, 8) != 0) throw SerialiseError("NotAnIndexFile", path);
    if (buf.size() < 12) throw SerialiseError("TruncatedIndexFile", path);
    uint32_t version;
    std::memcpy(&version, buf.data() + 8, 4);
    if (version != FORMAT_VERSION)
        throw SerialiseError("UnsupportedFormatVersion", "found " + std::to_string(version) + ", supported " + std::to_string(FORMAT_VERSION));
    if (buf.size() < 14) throw SerialiseError("TruncatedIndexFile", path);
    const uint8_t width = static_cast<uint8_t>(buf[12]), klen = static_cast<uint8_t>(buf[13]);
    if (buf.size() < 14u + klen) throw SerialiseError("TruncatedIndexFile", path);
    const std::string found = buf.substr(14, klen);
    if (found != kind) throw SerialiseError("IndexKindMismatch", "expected '" + kind + "', found '" + found + "'");
    if (width != float_width) throw SerialiseError("FloatWidthMismatch", "expected " + std::to_string(float_width) + ", found " + std::to_string(width));
    return 14u + klen;
}

inline std::string read_file(const std::string& dir, std::string* path_out) {
    const std::string path = dir + "/" + INDEX_FILE;
    *path_out = path;
    std::ifstream f(path, std::ios::binary);
    if (!f) throw SerialiseError("IoError", "cannot open " + path);
    return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
// save_index (mod.rs:256-296): temporary name, then rename into place.  The directory must exist.
inline void write_file(const std::string& dir, const std::string& bytes) {
    const std::string path = dir + "/" + INDEX_FILE, tmp = path + ".tmp";
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        if (!f) throw SerialiseError("IoError", "cannot create " + tmp);
        f.write(bytes.data(), static_cast<std::streamsize>(bytes.size()));
        if (!f) throw SerialiseError("IoError", "short write to " + tmp);
    }
    if (std::rename(tmp.c_str(), path.c_str()) != 0) throw SerialiseError("IoError", "cannot rename " + tmp);
}

}  // namespace detail

inline void save_exhaustive(const std::string& dir, const SavedExhaustive& ix) {
    std::string out = detail::header("exhaustive", 4);
    detail::put_f32_vec(out, ix.vectors_flat);
    detail::put_varint(out, ix.dim);
    detail::put_varint(out, ix.n);
    detail::put_f32_vec(out, ix.norms);
    detail::put_varint(out, ix.metric);
    detail::write_file(dir, out);
}

inline SavedExhaustive load_exhaustive(const std::string& dir) {
    std::string path;
    const std::string buf = detail::read_file(dir, &path);
    detail::Reader r{buf, detail::read_header(buf, path, "exhaustive", 4)};
    SavedExhaustive ix;
    ix.vectors_flat = r.f32_vec();
    ix.dim = r.varint();
    ix.n = r.varint();
    ix.norms = r.f32_vec();
    ix.metric = r.dist();
    if (r.pos != buf.size()) throw SerialiseError("TrailingBytes", path);   // mod.rs:316-322
    if (ix.vectors_flat.size() != ix.n * ix.dim) throw SerialiseError("DecodeError", "vector count does not match n * dim");
    return ix;
}

inline void save_ivf(const std::string& dir, const SavedIvf& ix) {
    std::string out = detail::header("ivf", 4);
    detail::put_f32_vec(out, ix.vectors_flat);
    detail::put_varint(out, ix.dim);
    detail::put_varint(out, ix.n);
    detail::put_f32_vec(out, ix.norms);
    detail::put_varint(out, ix.metric);
    detail::put_f32_vec(out, ix.centroids);
    detail::put_f32_vec(out, ix.centroids_norm);
    detail::put_usize_vec(out, ix.all_indices);
    detail::put_usize_vec(out, ix.offsets);
    detail::put_varint(out, ix.nlist);
    detail::put_usize_vec(out, ix.original_ids);
    detail::write_file(dir, out);
}

inline SavedIvf load_ivf(const std::string& dir) {
    std::string path;
    const std::string buf = detail::read_file(dir, &path);
    detail::Reader r{buf, detail::read_header(buf, path, "ivf", 4)};
    SavedIvf ix;
    ix.vectors_flat = r.f32_vec();
    ix.dim = r.varint();
    ix.n = r.varint();
    ix.norms = r.f32_vec();
    ix.metric = r.dist();
    ix.centroids = r.f32_vec();
    ix.centroids_norm = r.f32_vec();
    ix.all_indices = r.usize_vec();
    ix.offsets = r.usize_vec();
    ix.nlist = r.varint();
    ix.original_ids = r.usize_vec();
    if (r.pos != buf.size()) throw SerialiseError("TrailingBytes", path);
    if (ix.vectors_flat.size() != ix.n * ix.dim || ix.centroids.size() != ix.nlist * ix.dim || ix.offsets.size() != ix.nlist + 1)
        throw SerialiseError("DecodeError", "inconsistent sizes in the ivf payload");
    return ix;
}

}  // namespace annb200::serialise
