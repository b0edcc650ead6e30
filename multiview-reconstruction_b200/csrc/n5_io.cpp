// N5 datasets at the boundary of the loop: the PSF store of a SpimData project -- PointSpreadFunction.load / save
// (M/fiji/spimdata/pointspreadfunctions/PointSpreadFunction.java:119-162: dataset "psf_t<T>_v<V>" in <project>/psf.n5, float32, 128^3
// blocks, gzip level 1) -- and the N5 export of the deconvolved volume (M/process/export/ExportN5Api.java).  The reference goes through
// the n5 / n5-imglib2 libraries (third-party, absent from the tree); this restates the published N5 file-system format
// (github.com/saalfeldlab/n5, "File-system specification"):
//   <dataset>/attributes.json   {"dimensions":[x,y,z], "blockSize":[bx,by,bz], "dataType":"float32", "compression":{"type":"raw"|"gzip",...}}
//   <dataset>/<gx>/<gy>/<gz>    one file per block at grid position (gx, gy, gz); missing file = all zeros
//   block file, big endian:     uint16 mode (0 = default, 1 = varlength), uint16 ndim, ndim x uint32 actual block size,
//                               [mode 1: uint32 number of elements], payload = the block's elements, first dimension fastest, big endian,
//                               as they are ("raw") or as one gzip member ("gzip"; "useZlib": true -> a zlib stream)
// Host code only.  3-d datasets (trailing dimensions of size 1 are accepted), element types uint8 / int8 / uint16 / int16 / uint32 /
// int32 / float32 / float64 on input, float32 on output.
#include <sys/stat.h>
#include <sys/types.h>
#include <zlib.h>

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "engine.h"

namespace mvd {

namespace {
std::string slurp(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw Error("N5: cannot open " + path);
    std::string s;
    char buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, n);
    std::fclose(f);
    return s;
}
bool exists(const std::string& path) { struct stat st; return ::stat(path.c_str(), &st) == 0; }
void mkdirs(const std::string& path) {
    for (size_t i = 1; i <= path.size(); ++i)
        if (i == path.size() || path[i] == '/') {
            const std::string p = path.substr(0, i);
            if (::mkdir(p.c_str(), 0777) != 0 && errno != EEXIST) throw Error("N5: cannot create directory " + p);
        }
}
// the few attributes the format needs, out of a (flat or one level nested) JSON object
size_t find_key(const std::string& js, const std::string& key) {
    const std::string q = "\"" + key + "\"";
    const size_t k = js.find(q);
    if (k == std::string::npos) return k;
    const size_t c = js.find(':', k + q.size());
    return c == std::string::npos ? c : c + 1;
}
std::vector<long long> json_int_array(const std::string& js, const std::string& key) {
    std::vector<long long> v;
    size_t p = find_key(js, key);
    if (p == std::string::npos) return v;
    p = js.find('[', p);
    const size_t e = js.find(']', p);
    if (p == std::string::npos || e == std::string::npos) return v;
    const char* s = js.c_str() + p + 1;
    const char* end = js.c_str() + e;
    while (s < end) {
        char* next = nullptr;
        const long long x = std::strtoll(s, &next, 10);
        if (next == s) { ++s; continue; }
        v.push_back(x);
        s = next;
    }
    return v;
}
std::string json_string(const std::string& js, const std::string& key, size_t from = 0) {
    const std::string q = "\"" + key + "\"";
    const size_t k = js.find(q, from);
    if (k == std::string::npos) return "";
    const size_t c = js.find(':', k + q.size());
    const size_t a = js.find('"', c);
    const size_t b = a == std::string::npos ? a : js.find('"', a + 1);
    if (c == std::string::npos || b == std::string::npos) return "";
    return js.substr(a + 1, b - a - 1);
}

struct Attributes {
    long long dims[3] = {1, 1, 1};
    long long block[3] = {1, 1, 1};
    int ndim = 3;
    std::string dtype, compression;
};
Attributes read_attributes(const std::string& dir) {
    const std::string js = slurp(dir + "/attributes.json");
    Attributes a;
    const std::vector<long long> d = json_int_array(js, "dimensions"), b = json_int_array(js, "blockSize");
    if (d.empty() || d.size() != b.size()) throw Error("N5: attributes.json without matching dimensions / blockSize");
    for (size_t i = 3; i < d.size(); ++i)
        if (d[i] != 1) throw Error("N5: only 3-d datasets (trailing dimensions of size 1) are supported");
    a.ndim = (int)d.size();
    for (size_t i = 0; i < d.size() && i < 3; ++i) { a.dims[i] = d[i]; a.block[i] = b[i]; }
    for (int i = 0; i < 3; ++i)
        if (a.dims[i] < 1 || a.block[i] < 1 || a.dims[i] > 0x7fffffff) throw Error("N5: bad dimensions");
    a.dtype = json_string(js, "dataType");
    const size_t c = js.find("\"compression\"");
    if (c != std::string::npos) {
        a.compression = json_string(js, "type", c);
        if (a.compression.empty()) a.compression = json_string(js, "compression");         // (string-valued in very old containers)
    } else {
        a.compression = json_string(js, "compressionType");                                 // N5 1.x
    }
    if (a.compression.empty()) a.compression = "raw";
    return a;
}
size_t dtype_size(const std::string& t) {
    if (t == "uint8" || t == "int8") return 1;
    if (t == "uint16" || t == "int16") return 2;
    if (t == "uint32" || t == "int32" || t == "float32") return 4;
    if (t == "float64") return 8;
    throw Error("N5: unsupported dataType '" + t + "'");
}
float element(const unsigned char* p, const std::string& t) {            // one big-endian element -> float
    if (t == "uint8") return (float)p[0];
    if (t == "int8") return (float)(signed char)p[0];
    if (t == "uint16") return (float)(uint16_t)(p[0] << 8 | p[1]);
    if (t == "int16") return (float)(int16_t)(uint16_t)(p[0] << 8 | p[1]);
    const uint32_t u = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
    if (t == "uint32") return (float)u;
    if (t == "int32") return (float)(int32_t)u;
    if (t == "float32") { float f; std::memcpy(&f, &u, 4); return f; }
    const uint64_t w = (uint64_t)u << 32 | ((uint64_t)p[4] << 24 | (uint64_t)p[5] << 16 | (uint64_t)p[6] << 8 | p[7]);
    double d; std::memcpy(&d, &w, 8); return (float)d;
}
std::vector<unsigned char> inflate_all(const unsigned char* src, size_t n, size_t expect) {
    std::vector<unsigned char> out(expect);
    z_stream z;
    std::memset(&z, 0, sizeof(z));
    if (inflateInit2(&z, 15 + 32) != Z_OK) throw Error("N5: zlib initialisation failed");      // gzip or zlib header, detected
    z.next_in = const_cast<unsigned char*>(src); z.avail_in = (uInt)n;
    z.next_out = out.data(); z.avail_out = (uInt)expect;
    const int rc = inflate(&z, Z_FINISH);
    const size_t got = z.total_out;
    inflateEnd(&z);
    if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || got != expect) throw Error("N5: corrupt compressed block");
    return out;
}
std::vector<unsigned char> gzip_all(const unsigned char* src, size_t n, int level) {
    z_stream z;
    std::memset(&z, 0, sizeof(z));
    if (deflateInit2(&z, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw Error("N5: zlib initialisation failed");
    std::vector<unsigned char> out(deflateBound(&z, (uLong)n) + 32);
    z.next_in = const_cast<unsigned char*>(src); z.avail_in = (uInt)n;
    z.next_out = out.data(); z.avail_out = (uInt)out.size();
    const int rc = deflate(&z, Z_FINISH);
    out.resize(z.total_out);
    deflateEnd(&z);
    if (rc != Z_STREAM_END) throw Error("N5: compression failed");
    return out;
}
void put_u16(std::vector<unsigned char>& b, unsigned v) { b.push_back((unsigned char)(v >> 8)); b.push_back((unsigned char)v); }
void put_u32(std::vector<unsigned char>& b, uint32_t v) { for (int s = 24; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
}  // namespace

void n5_dims(const char* dataset_dir, int dims[3]) {
    const Attributes a = read_attributes(dataset_dir);
    for (int i = 0; i < 3; ++i) dims[i] = (int)a.dims[i];
}

std::vector<float> n5_read_f32(const char* dataset_dir, int dims[3]) {
    const std::string dir = dataset_dir;
    const Attributes a = read_attributes(dir);
    const size_t es = dtype_size(a.dtype);
    if (a.compression != "raw" && a.compression != "gzip") throw Error("N5: unsupported compression '" + a.compression + "' (raw and gzip are)");
    for (int i = 0; i < 3; ++i) dims[i] = (int)a.dims[i];
    std::vector<float> out((size_t)a.dims[0] * a.dims[1] * a.dims[2], 0.f);
    long long grid[3];
    for (int i = 0; i < 3; ++i) grid[i] = (a.dims[i] + a.block[i] - 1) / a.block[i];
    for (long long gz = 0; gz < grid[2]; ++gz)
        for (long long gy = 0; gy < grid[1]; ++gy)
            for (long long gx = 0; gx < grid[0]; ++gx) {
                std::string path = dir + "/" + std::to_string(gx) + "/" + std::to_string(gy) + "/" + std::to_string(gz);
                for (int i = 3; i < a.ndim; ++i) path += "/0";
                if (!exists(path)) continue;                                   // missing block = zeros
                const std::string raw = slurp(path);
                const unsigned char* p = reinterpret_cast<const unsigned char*>(raw.data());
                if (raw.size() < 4) throw Error("N5: truncated block " + path);
                const unsigned mode = p[0] << 8 | p[1], nd = p[2] << 8 | p[3];
                if (mode > 1 || nd != (unsigned)a.ndim || raw.size() < 4 + 4 * (size_t)nd + (mode == 1 ? 4 : 0)) throw Error("N5: bad block header in " + path);
                long long bs[3] = {1, 1, 1};
                size_t count = 1;
                for (unsigned i = 0; i < nd; ++i) {
                    const unsigned char* q = p + 4 + 4 * i;
                    const uint32_t v = (uint32_t)q[0] << 24 | (uint32_t)q[1] << 16 | (uint32_t)q[2] << 8 | q[3];
                    if (i < 3) bs[i] = v; else if (v != 1) throw Error("N5: only 3-d blocks are supported");
                    count *= v;
                }
                const size_t hdr = 4 + 4 * (size_t)nd + (mode == 1 ? 4 : 0);
                const long long x0 = gx * a.block[0], y0 = gy * a.block[1], z0 = gz * a.block[2];
                if (bs[0] > a.block[0] || bs[1] > a.block[1] || bs[2] > a.block[2]) throw Error("N5: block larger than blockSize in " + path);
                std::vector<unsigned char> plain;
                const unsigned char* data;
                if (a.compression == "gzip") { plain = inflate_all(p + hdr, raw.size() - hdr, count * es); data = plain.data(); }
                else { if (raw.size() - hdr < count * es) throw Error("N5: truncated block " + path); data = p + hdr; }
                for (long long z = 0; z < bs[2] && z0 + z < a.dims[2]; ++z)
                    for (long long y = 0; y < bs[1] && y0 + y < a.dims[1]; ++y) {
                        const unsigned char* row = data + ((size_t)(z * bs[1] + y) * bs[0]) * es;
                        float* o = out.data() + ((size_t)(z0 + z) * a.dims[1] + (y0 + y)) * a.dims[0] + x0;
                        for (long long x = 0; x < bs[0] && x0 + x < a.dims[0]; ++x) o[x] = element(row + x * es, a.dtype);
                    }
            }
    return out;
}

void n5_write_f32(const char* dataset_dir, const float* data, const int dims[3], const int block[3], int gzip_level) {
    const std::string dir = dataset_dir;
    for (int i = 0; i < 3; ++i)
        if (dims[i] < 1 || block[i] < 1) throw Error("N5: bad dimensions / blockSize");
    if (gzip_level > 9) gzip_level = 9;
    mkdirs(dir);
    {
        std::string js = "{\"dataType\":\"float32\",\"compression\":";
        if (gzip_level < 0) js += "{\"type\":\"raw\"}";
        else js += "{\"type\":\"gzip\",\"useZlib\":false,\"level\":" + std::to_string(gzip_level) + "}";
        js += ",\"blockSize\":[" + std::to_string(block[0]) + "," + std::to_string(block[1]) + "," + std::to_string(block[2]) + "]";
        js += ",\"dimensions\":[" + std::to_string(dims[0]) + "," + std::to_string(dims[1]) + "," + std::to_string(dims[2]) + "]}";
        FILE* f = std::fopen((dir + "/attributes.json").c_str(), "wb");
        if (!f) throw Error("N5: cannot write " + dir + "/attributes.json");
        std::fwrite(js.data(), 1, js.size(), f);
        std::fclose(f);
    }
    long long grid[3];
    for (int i = 0; i < 3; ++i) grid[i] = ((long long)dims[i] + block[i] - 1) / block[i];
    std::vector<unsigned char> buf;
    for (long long gz = 0; gz < grid[2]; ++gz)
        for (long long gy = 0; gy < grid[1]; ++gy)
            for (long long gx = 0; gx < grid[0]; ++gx) {
                const long long x0 = gx * block[0], y0 = gy * block[1], z0 = gz * block[2];
                const long long bx = std::min<long long>(block[0], dims[0] - x0), by = std::min<long long>(block[1], dims[1] - y0),
                                bz = std::min<long long>(block[2], dims[2] - z0);
                buf.clear();
                buf.reserve((size_t)(bx * by * bz) * 4);
                for (long long z = 0; z < bz; ++z)
                    for (long long y = 0; y < by; ++y) {
                        const float* row = data + ((size_t)(z0 + z) * dims[1] + (y0 + y)) * dims[0] + x0;
                        for (long long x = 0; x < bx; ++x) { uint32_t u; std::memcpy(&u, row + x, 4); put_u32(buf, u); }
                    }
                std::vector<unsigned char> file;
                put_u16(file, 0); put_u16(file, 3);
                put_u32(file, (uint32_t)bx); put_u32(file, (uint32_t)by); put_u32(file, (uint32_t)bz);
                if (gzip_level < 0) file.insert(file.end(), buf.begin(), buf.end());
                else { const std::vector<unsigned char> c = gzip_all(buf.data(), buf.size(), gzip_level); file.insert(file.end(), c.begin(), c.end()); }
                const std::string sub = dir + "/" + std::to_string(gx) + "/" + std::to_string(gy);
                mkdirs(sub);
                FILE* f = std::fopen((sub + "/" + std::to_string(gz)).c_str(), "wb");
                if (!f) throw Error("N5: cannot write a block under " + sub);
                const size_t w = std::fwrite(file.data(), 1, file.size(), f);
                std::fclose(f);
                if (w != file.size()) throw Error("N5: short write under " + sub);
            }
}

// OME-Zarr (NGFF 0.4) export of the result, the second container format of ExportN5Api (M/process/export/ExportN5Api.java; n5-zarr is
// third-party): a Zarr v2 group with one multiscale level,
//   <path>/.zgroup   {"zarr_format":2}
//   <path>/.zattrs   {"multiscales":[{"version":"0.4","axes":[z,y,x (space, micrometer)],"datasets":[{"path":"0","coordinateTransformations":[scale]}]}]}
//   <path>/0/.zarray {"zarr_format":2,"shape":[z,y,x],"chunks":[cz,cy,cx],"dtype":"<f4","order":"C","fill_value":0,"dimension_separator":"/",
//                     "compressor":null | {"id":"gzip","level":L},"filters":null}
//   <path>/0/<iz>/<iy>/<ix>   every chunk in full chunk size (edge chunks padded with the fill value), C order, little endian, gzip'ed or raw
void zarr_write_f32(const char* path, const float* data, const int dims[3], const int chunk[3], int gzip_level, const double* voxel_size) {
    const std::string dir = path;
    for (int i = 0; i < 3; ++i)
        if (dims[i] < 1 || chunk[i] < 1) throw Error("Zarr: bad dimensions / chunk size");
    if (gzip_level > 9) gzip_level = 9;
    const std::string arr = dir + "/0";
    mkdirs(arr);
    auto put_file = [](const std::string& fn, const std::string& text) {
        FILE* f = std::fopen(fn.c_str(), "wb");
        if (!f) throw Error("Zarr: cannot write " + fn);
        std::fwrite(text.data(), 1, text.size(), f);
        std::fclose(f);
    };
    const double vs[3] = {voxel_size ? voxel_size[0] : 1.0, voxel_size ? voxel_size[1] : 1.0, voxel_size ? voxel_size[2] : 1.0};
    put_file(dir + "/.zgroup", "{\"zarr_format\":2}");
    char scale[160];
    std::snprintf(scale, sizeof(scale), "[%.17g,%.17g,%.17g]", vs[2], vs[1], vs[0]);
    put_file(dir + "/.zattrs",
             std::string("{\"multiscales\":[{\"version\":\"0.4\",\"name\":\"deconvolved\",\"axes\":["
                         "{\"name\":\"z\",\"type\":\"space\",\"unit\":\"micrometer\"},{\"name\":\"y\",\"type\":\"space\",\"unit\":\"micrometer\"},"
                         "{\"name\":\"x\",\"type\":\"space\",\"unit\":\"micrometer\"}],"
                         "\"datasets\":[{\"path\":\"0\",\"coordinateTransformations\":[{\"type\":\"scale\",\"scale\":") + scale + "}]}]}]}");
    std::string za = "{\"zarr_format\":2,\"shape\":[" + std::to_string(dims[2]) + "," + std::to_string(dims[1]) + "," + std::to_string(dims[0]) + "],\"chunks\":[" +
                     std::to_string(chunk[2]) + "," + std::to_string(chunk[1]) + "," + std::to_string(chunk[0]) + "],\"dtype\":\"<f4\",\"order\":\"C\",\"fill_value\":0,"
                     "\"dimension_separator\":\"/\",\"filters\":null,\"compressor\":";
    za += gzip_level < 0 ? std::string("null}") : "{\"id\":\"gzip\",\"level\":" + std::to_string(gzip_level) + "}}";
    put_file(arr + "/.zarray", za);
    long long grid[3];
    for (int i = 0; i < 3; ++i) grid[i] = ((long long)dims[i] + chunk[i] - 1) / chunk[i];
    std::vector<float> buf((size_t)chunk[0] * chunk[1] * chunk[2]);
    for (long long gz = 0; gz < grid[2]; ++gz)
        for (long long gy = 0; gy < grid[1]; ++gy)
            for (long long gx = 0; gx < grid[0]; ++gx) {
                std::fill(buf.begin(), buf.end(), 0.f);
                const long long x0 = gx * chunk[0], y0 = gy * chunk[1], z0 = gz * chunk[2];
                const long long bx = std::min<long long>(chunk[0], dims[0] - x0);
                for (long long z = 0; z < chunk[2] && z0 + z < dims[2]; ++z)
                    for (long long y = 0; y < chunk[1] && y0 + y < dims[1]; ++y)
                        std::memcpy(buf.data() + ((size_t)z * chunk[1] + y) * chunk[0], data + ((size_t)(z0 + z) * dims[1] + (y0 + y)) * dims[0] + x0,
                                    sizeof(float) * (size_t)bx);                     // (x86 / CUDA hosts are little endian: "<f4" as stored)
                const unsigned char* raw = reinterpret_cast<const unsigned char*>(buf.data());
                const size_t nbytes = buf.size() * sizeof(float);
                const std::string sub = arr + "/" + std::to_string(gz) + "/" + std::to_string(gy);
                mkdirs(sub);
                FILE* f = std::fopen((sub + "/" + std::to_string(gx)).c_str(), "wb");
                if (!f) throw Error("Zarr: cannot write a chunk under " + sub);
                size_t w, want;
                if (gzip_level < 0) { want = nbytes; w = std::fwrite(raw, 1, nbytes, f); }
                else { const std::vector<unsigned char> c = gzip_all(raw, nbytes, gzip_level); want = c.size(); w = std::fwrite(c.data(), 1, c.size(), f); }
                std::fclose(f);
                if (w != want) throw Error("Zarr: short write under " + sub);
            }
}

}  // namespace mvd
