// Minimal TIFF stack IO at the boundary of the loop: PsiInitFromFile opens a start image "as 32 bit"
// (M/process/deconvolution/init/PsiInitFromFile.java:66-93, IOFunctions.openAs32Bit) and the result is saved as a 3-d TIFF
// (M/process/export/Save3dTIFF.java).  The reference goes through ImageJ's opener / FileSaver (third-party); this restates the subset
// of baseline TIFF 6.0 that ImageJ produces for stacks: uncompressed, one sample per pixel, 8/16/32-bit integer or 32-bit float, strips,
// either one IFD per slice or -- for stacks beyond 4 GB -- ImageJ's single IFD followed by contiguous slices ("images=N" in the
// description).  Host code only.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "engine.h"

namespace mvd {

namespace {
struct File {
    FILE* f = nullptr;
    explicit File(const char* path, const char* mode) : f(std::fopen(path, mode)) {}
    ~File() { if (f) std::fclose(f); }
    File(const File&) = delete;
    File& operator=(const File&) = delete;
};
struct Reader {
    FILE* f; bool big;
    void seek(uint64_t o) const { if (fseeko(f, (off_t)o, SEEK_SET) != 0) throw Error("TIFF: seek failed"); }
    void bytes(void* p, size_t n) const { if (std::fread(p, 1, n, f) != n) throw Error("TIFF: unexpected end of file"); }
    uint16_t u16() const { unsigned char b[2]; bytes(b, 2); return big ? (uint16_t)(b[0] << 8 | b[1]) : (uint16_t)(b[1] << 8 | b[0]); }
    uint32_t u32() const {
        unsigned char b[4]; bytes(b, 4);
        return big ? ((uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3]) : ((uint32_t)b[3] << 24 | (uint32_t)b[2] << 16 | (uint32_t)b[1] << 8 | b[0]);
    }
};
struct Ifd {
    uint32_t width = 0, height = 0, bits = 1, compression = 1, samples = 1, format = 1, rows_per_strip = 0xFFFFFFFFu;
    std::vector<uint64_t> offsets, counts;
    std::string description;
    uint32_t next = 0;
};
size_t type_size(uint16_t t) { return t == 1 || t == 2 || t == 6 || t == 7 ? 1 : t == 3 || t == 8 ? 2 : t == 4 || t == 9 || t == 11 ? 4 : 8; }

Ifd read_ifd(const Reader& r, uint64_t at) {
    Ifd d;
    r.seek(at);
    const uint16_t n = r.u16();
    struct Entry { uint16_t tag, type; uint32_t count; long pos; };
    std::vector<Entry> es(n);
    for (Entry& e : es) {
        e.tag = r.u16(); e.type = r.u16(); e.count = r.u32(); e.pos = std::ftell(r.f);
        unsigned char skip[4]; r.bytes(skip, 4);
    }
    d.next = r.u32();
    auto values = [&](const Entry& e) {
        std::vector<uint64_t> v(e.count);
        const size_t total = type_size(e.type) * e.count;
        r.seek((uint64_t)e.pos);
        if (total > 4) { const uint32_t off = r.u32(); r.seek(off); }
        for (uint32_t i = 0; i < e.count; ++i) {
            if (e.type == 3) v[i] = r.u16();
            else if (e.type == 4) v[i] = r.u32();
            else if (e.type == 1) { unsigned char b; r.bytes(&b, 1); v[i] = b; }
            else throw Error("TIFF: unsupported field type");
        }
        return v;
    };
    for (const Entry& e : es) {
        switch (e.tag) {
            case 256: d.width = (uint32_t)values(e).at(0); break;
            case 257: d.height = (uint32_t)values(e).at(0); break;
            case 258: d.bits = (uint32_t)values(e).at(0); break;
            case 259: d.compression = (uint32_t)values(e).at(0); break;
            case 277: d.samples = (uint32_t)values(e).at(0); break;
            case 278: d.rows_per_strip = (uint32_t)values(e).at(0); break;
            case 339: d.format = (uint32_t)values(e).at(0); break;
            case 273: d.offsets = values(e); break;
            case 279: d.counts = values(e); break;
            case 270: {
                std::string s(e.count, '\0');
                r.seek((uint64_t)e.pos);
                if (e.count > 4) { const uint32_t off = r.u32(); r.seek(off); }
                r.bytes(&s[0], e.count);
                d.description = s;
                break;
            }
            default: break;
        }
    }
    if (d.width == 0 || d.height == 0 || d.offsets.empty()) throw Error("TIFF: incomplete image directory");
    if (d.compression != 1) throw Error("TIFF: compressed files are not supported (save the start image uncompressed)");
    if (d.samples != 1) throw Error("TIFF: only single-channel images are supported");
    if (!((d.bits == 8 || d.bits == 16 || d.bits == 32) && (d.format == 1 || d.format == 2 || (d.format == 3 && d.bits == 32))))
        throw Error("TIFF: unsupported sample type");
    return d;
}

void convert(const unsigned char* raw, size_t n, const Ifd& d, bool big, float* out) {       // "openAs32Bit"
    const int bytes = (int)d.bits / 8;
    for (size_t i = 0; i < n; ++i) {
        const unsigned char* p = raw + i * bytes;
        uint32_t v = 0;
        for (int b = 0; b < bytes; ++b) v |= (uint32_t)p[big ? bytes - 1 - b : b] << (8 * b);
        if (d.format == 3) { float f; std::memcpy(&f, &v, 4); out[i] = f; }
        else if (d.format == 2) out[i] = bytes == 1 ? (float)(int8_t)v : bytes == 2 ? (float)(int16_t)v : (float)(int32_t)v;
        else out[i] = (float)v;
    }
}

struct Stack { bool big; std::vector<Ifd> ifds; uint32_t slices; };
Stack open_stack(FILE* f) {
    unsigned char h[4];
    Reader r{f, false};
    r.bytes(h, 4);
    if (h[0] == 'I' && h[1] == 'I') r.big = false;
    else if (h[0] == 'M' && h[1] == 'M') r.big = true;
    else throw Error("not a TIFF file");
    if ((r.big ? h[3] : h[2]) != 42) throw Error("TIFF: bad magic (BigTIFF is not supported)");
    Stack s;
    s.big = r.big;
    uint32_t at = r.u32();
    while (at != 0 && s.ifds.size() < (1u << 20)) { s.ifds.push_back(read_ifd(r, at)); at = s.ifds.back().next; }
    if (s.ifds.empty()) throw Error("TIFF: no image");
    s.slices = (uint32_t)s.ifds.size();
    if (s.ifds.size() == 1) {                                   // ImageJ stack with a single directory: "images=N"
        const size_t p = s.ifds[0].description.find("images=");
        if (p != std::string::npos) { const long n = std::atol(s.ifds[0].description.c_str() + p + 7); if (n > 1) s.slices = (uint32_t)n; }
    }
    for (const Ifd& d : s.ifds)
        if (d.width != s.ifds[0].width || d.height != s.ifds[0].height || d.bits != s.ifds[0].bits || d.format != s.ifds[0].format)
            throw Error("TIFF: slices of different size or type");
    return s;
}
}  // namespace

void tiff_dims(const char* path, int dims[3]) {
    File fh(path, "rb");
    if (!fh.f) throw Error(std::string("cannot open ") + path);
    const Stack s = open_stack(fh.f);
    dims[0] = (int)s.ifds[0].width; dims[1] = (int)s.ifds[0].height; dims[2] = (int)s.slices;
}

std::vector<float> tiff_read_f32(const char* path, int dims[3]) {
    File fh(path, "rb");
    if (!fh.f) throw Error(std::string("cannot open ") + path);
    const Stack s = open_stack(fh.f);
    const Reader r{fh.f, s.big};
    const Ifd& d0 = s.ifds[0];
    dims[0] = (int)d0.width; dims[1] = (int)d0.height; dims[2] = (int)s.slices;
    const size_t plane = (size_t)d0.width * d0.height, bytes = d0.bits / 8;
    std::vector<float> out(plane * s.slices);
    std::vector<unsigned char> raw(plane * bytes);
    for (uint32_t z = 0; z < s.slices; ++z) {
        if (s.ifds.size() == s.slices) {
            const Ifd& d = s.ifds[z];
            size_t got = 0;
            for (size_t i = 0; i < d.offsets.size() && got < raw.size(); ++i) {
                size_t n = i < d.counts.size() ? (size_t)d.counts[i] : raw.size() - got;
                if (n > raw.size() - got) n = raw.size() - got;
                r.seek(d.offsets[i]);
                r.bytes(raw.data() + got, n);
                got += n;
            }
            if (got != raw.size()) throw Error("TIFF: strip data shorter than the image");
        } else {                                                // ImageJ: contiguous slices after the first strip offset
            r.seek(d0.offsets[0] + (uint64_t)z * raw.size());
            r.bytes(raw.data(), raw.size());
        }
        convert(raw.data(), plane, d0, s.big, out.data() + (size_t)z * plane);
    }
    return out;
}

// little-endian 32-bit float stack with an ImageJ description; one directory per slice below 4 GB, ImageJ's contiguous layout above
void tiff_write_f32(const char* path, const float* data, const int dims[3]) {
    if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1) throw Error("TIFF: empty image");
    File fh(path, "wb");
    if (!fh.f) throw Error(std::string("cannot create ") + path);
    const uint64_t plane_bytes = (uint64_t)dims[0] * dims[1] * 4, total = plane_bytes * (uint64_t)dims[2];
    const std::string desc = "ImageJ=1.53t\nimages=" + std::to_string(dims[2]) + "\nslices=" + std::to_string(dims[2]) + "\n";
    const uint32_t desc_len = (uint32_t)desc.size() + 1;
    const bool per_slice = total + 8 + desc_len + (uint64_t)dims[2] * 200 < 0xFFFF0000ull;
    if (plane_bytes > 0xFFFF0000ull) throw Error("TIFF: a single slice beyond 4 GB is not supported");
    auto w16 = [&](uint16_t v) { unsigned char b[2] = {(unsigned char)(v & 255), (unsigned char)(v >> 8)}; std::fwrite(b, 1, 2, fh.f); };
    auto w32 = [&](uint32_t v) { unsigned char b[4] = {(unsigned char)(v & 255), (unsigned char)(v >> 8 & 255), (unsigned char)(v >> 16 & 255), (unsigned char)(v >> 24)}; std::fwrite(b, 1, 4, fh.f); };
    auto entry = [&](uint16_t tag, uint16_t type, uint32_t count, uint32_t value) { w16(tag); w16(type); w32(count); if (type == 3 && count == 1) { w16((uint16_t)value); w16(0); } else w32(value); };
    const uint32_t n_entries = 11, ifd_bytes = 2 + 12 * n_entries + 4;
    const uint32_t n_ifds = per_slice ? (uint32_t)dims[2] : 1;
    // layout: header | description | directories | pixel data
    const uint32_t desc_at = 8, ifd_at = desc_at + ((desc_len + 1) & ~1u);
    const uint64_t data_at = (uint64_t)ifd_at + (uint64_t)ifd_bytes * n_ifds;
    std::fwrite("II", 1, 2, fh.f); w16(42); w32(ifd_at);
    std::fwrite(desc.c_str(), 1, desc_len, fh.f);
    if (desc_len & 1) std::fputc(0, fh.f);
    for (uint32_t i = 0; i < n_ifds; ++i) {
        w16((uint16_t)n_entries);
        entry(256, 4, 1, (uint32_t)dims[0]);
        entry(257, 4, 1, (uint32_t)dims[1]);
        entry(258, 3, 1, 32);
        entry(259, 3, 1, 1);
        entry(262, 3, 1, 1);                                                // BlackIsZero
        entry(270, 2, desc_len, desc_at);
        entry(273, 4, 1, (uint32_t)(data_at + (uint64_t)i * plane_bytes));
        entry(277, 3, 1, 1);
        entry(278, 4, 1, (uint32_t)dims[1]);
        entry(279, 4, 1, (uint32_t)plane_bytes);
        entry(339, 3, 1, 3);                                                // IEEE float
        w32(i + 1 < n_ifds ? ifd_at + ifd_bytes * (i + 1) : 0);
    }
    const size_t n = (size_t)dims[0] * dims[1] * dims[2];
    const uint16_t probe = 1;
    if (*reinterpret_cast<const unsigned char*>(&probe) == 1) {             // little-endian host: the floats go out as they are
        if (std::fwrite(data, 4, n, fh.f) != n) throw Error("TIFF: write failed");
    } else {
        for (size_t i = 0; i < n; ++i) { uint32_t v; std::memcpy(&v, data + i, 4); w32(v); }
    }
    if (std::fflush(fh.f) != 0) throw Error("TIFF: write failed");
}

}  // namespace mvd
