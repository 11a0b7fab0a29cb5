// dgd_dump.h — "DGD1" named-array container used to move stage outputs between
// the reference hook (oracle/ref_hook.cpp), the host glue library, tests and
// bench.py.  A file is the 4-byte magic "DGD1" followed by records:
//   u32 name_len | name bytes | u32 dtype | u64 count | raw little-endian data
// dtype: 0=u8 1=i32 2=u32 3=i64 4=u64 5=f64.  Python reader: dipgenie_b200/dgd.py.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace dgd {

enum DType : uint32_t { U8 = 0, I32 = 1, U32 = 2, I64 = 3, U64 = 4, F64 = 5 };

inline size_t dtype_size(uint32_t t) {
    switch (t) { case U8: return 1; case I32: case U32: return 4; default: return 8; }
}

class Writer {
public:
    explicit Writer(const std::string& path) {
        fp_ = std::fopen(path.c_str(), "wb");
        if (fp_) std::fwrite("DGD1", 1, 4, fp_);
    }
    ~Writer() { if (fp_) std::fclose(fp_); }
    bool ok() const { return fp_ != nullptr; }

    void raw(const std::string& name, uint32_t dtype, const void* data, uint64_t count) {
        if (!fp_) return;
        uint32_t nl = (uint32_t)name.size();
        std::fwrite(&nl, 4, 1, fp_);
        std::fwrite(name.data(), 1, nl, fp_);
        std::fwrite(&dtype, 4, 1, fp_);
        std::fwrite(&count, 8, 1, fp_);
        if (count) std::fwrite(data, dtype_size(dtype), count, fp_);
    }
    void put(const std::string& n, const std::vector<uint8_t>& v)  { raw(n, U8,  v.data(), v.size()); }
    void put(const std::string& n, const std::vector<int32_t>& v)  { raw(n, I32, v.data(), v.size()); }
    void put(const std::string& n, const std::vector<uint32_t>& v) { raw(n, U32, v.data(), v.size()); }
    void put(const std::string& n, const std::vector<int64_t>& v)  { raw(n, I64, v.data(), v.size()); }
    void put(const std::string& n, const std::vector<uint64_t>& v) { raw(n, U64, v.data(), v.size()); }
    void put(const std::string& n, const std::vector<double>& v)   { raw(n, F64, v.data(), v.size()); }
    void put_str(const std::string& n, const std::string& s)       { raw(n, U8, s.data(), s.size()); }
    void put_i64(const std::string& n, int64_t x)                  { raw(n, I64, &x, 1); }

    // ragged list-of-lists as <name>.off (i64, size n+1) + <name>.val
    template <class Outer>
    void put_ragged_i32(const std::string& n, const Outer& lists) {
        std::vector<int64_t> off; off.reserve(lists.size() + 1); off.push_back(0);
        std::vector<int32_t> val;
        for (const auto& l : lists) {
            for (auto x : l) val.push_back((int32_t)x);
            off.push_back((int64_t)val.size());
        }
        put(n + ".off", off);
        put(n + ".val", val);
    }

private:
    std::FILE* fp_ = nullptr;
};

// Minimal reader: loads the whole file, hands out typed views by name.
class Reader {
public:
    explicit Reader(const std::string& path) {
        std::FILE* fp = std::fopen(path.c_str(), "rb");
        if (!fp) return;
        std::fseek(fp, 0, SEEK_END);
        long n = std::ftell(fp);
        std::fseek(fp, 0, SEEK_SET);
        buf_.resize((size_t)n);
        size_t got = n ? std::fread(buf_.data(), 1, (size_t)n, fp) : 0;
        std::fclose(fp);
        if (got != (size_t)n || n < 4 || std::memcmp(buf_.data(), "DGD1", 4) != 0) { buf_.clear(); return; }
        size_t pos = 4;
        while (pos + 16 <= buf_.size()) {
            uint32_t nl; std::memcpy(&nl, &buf_[pos], 4); pos += 4;
            if (pos + nl + 12 > buf_.size()) break;
            Entry e; e.name.assign((const char*)&buf_[pos], nl); pos += nl;
            std::memcpy(&e.dtype, &buf_[pos], 4); pos += 4;
            std::memcpy(&e.count, &buf_[pos], 8); pos += 8;
            e.off = pos;
            pos += (size_t)e.count * dtype_size(e.dtype);
            if (pos > buf_.size()) break;
            entries_.push_back(e);
        }
        ok_ = true;
    }
    bool ok() const { return ok_; }
    // Copies entry `name` into `out` converting to T; returns false if absent.
    template <class T>
    bool get(const std::string& name, std::vector<T>& out) const {
        for (const auto& e : entries_) {
            if (e.name != name) continue;
            out.resize((size_t)e.count);
            const uint8_t* p = &buf_[e.off];
            for (uint64_t i = 0; i < e.count; ++i) {
                switch (e.dtype) {
                    case U8:  out[i] = (T)p[i]; break;
                    case I32: { int32_t x; std::memcpy(&x, p + 4 * i, 4); out[i] = (T)x; break; }
                    case U32: { uint32_t x; std::memcpy(&x, p + 4 * i, 4); out[i] = (T)x; break; }
                    case I64: { int64_t x; std::memcpy(&x, p + 8 * i, 8); out[i] = (T)x; break; }
                    case U64: { uint64_t x; std::memcpy(&x, p + 8 * i, 8); out[i] = (T)x; break; }
                    default:  { double x; std::memcpy(&x, p + 8 * i, 8); out[i] = (T)x; break; }
                }
            }
            return true;
        }
        return false;
    }

private:
    struct Entry { std::string name; uint32_t dtype = 0; uint64_t count = 0; size_t off = 0; };
    std::vector<uint8_t> buf_;
    std::vector<Entry> entries_;
    bool ok_ = false;
};

}  // namespace dgd
