// az_ckpt.cpp — TensorFlow V2 checkpoint bundles on the host (no CUDA, no TensorFlow).
//
// What AlphaZeroNN::saveCheckpoint / loadCheckpoint exchange through the graph's save/restore ops
// (neural_network/alphazero_nn.cpp:189-214; tf.train.Saver V2 in python/src/build_graph.py:109): a pair of files
//   <prefix>.index                 an SSTable (TensorFlow's copy of the LevelDB table format, tensorflow/core/lib/io/table*):
//                                  key "" -> BundleHeaderProto, key <tensor name> -> BundleEntryProto, keys ascending
//   <prefix>.data-00000-of-00001   the raw little-endian tensor bytes, back to back, at the offsets the entries give
// (tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc}, tensorflow/core/protobuf/tensor_bundle.proto).  TensorFlow is an un-vendored,
// un-pinned dependency of the reference and is absent here, so this is a restatement of the published format; the reference ships no
// checkpoint to pin it against ("parity unpinned").  The tensor inventory a checkpoint of the shipped graph holds (163
// names: variables, BN moving statistics, the Adam slots "<var>/optimize", "<var>/optimize_1", beta1_power, beta2_power) is taken from
// the GraphDef's save/SaveV2/tensor_names and checked in tests/test_ckpt_cpu.py.
//
// Table format: data blocks | meta-index block | index block | 48-byte footer.
//   block  = entries, uint32 restart offsets, uint32 restart count; followed on disk by a 5-byte trailer: type (0 = uncompressed)
//            and the masked CRC32C of contents + type
//   entry  = varint32 shared key bytes, varint32 unshared key bytes, varint32 value bytes, key suffix, value
//   index  = one entry per data block: key >= every key of the block, value = BlockHandle (varint64 offset, varint64 size)
//   footer = meta-index handle, index handle, zero padding to 40 bytes, magic 0xdb4775248b80fb57 (little-endian)
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <new>

#include "az_b200.h"

void az_set_error(const char* fmt, ...);

namespace {

const uint64_t kTableMagic = 0xdb4775248b80fb57ull;
const size_t kBlockSize = 262144;        // table::Options::block_size
const int kRestartInterval = 16;         // table::Options::block_restart_interval

// ---- CRC32C (Castagnoli, reflected polynomial 0x82f63b78), masked as in tensorflow/core/lib/hash/crc32c.h
struct Crc32cTable {
    uint32_t t[256];
    Crc32cTable() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82f63b78u : c >> 1;
            t[i] = c;
        }
    }
};
uint32_t crc32c_extend(uint32_t crc, const void* data, size_t n)
{
    static const Crc32cTable T;
    const uint8_t* p = (const uint8_t*)data;
    uint32_t c = crc ^ 0xffffffffu;
    for (size_t i = 0; i < n; ++i) c = T.t[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c ^ 0xffffffffu;
}
const uint32_t kMaskDelta = 0xa282ead8u;
uint32_t crc_mask(uint32_t c) { return ((c >> 15) | (c << 17)) + kMaskDelta; }
uint32_t crc_unmask(uint32_t m) { uint32_t r = m - kMaskDelta; return (r >> 17) | (r << 15); }

// ---- varints / little-endian
void put_varint(std::string& s, uint64_t v) { while (v >= 128) { s.push_back((char)(v | 128)); v >>= 7; } s.push_back((char)v); }
void put_fixed32(std::string& s, uint32_t v) { for (int i = 0; i < 4; ++i) s.push_back((char)(v >> (8 * i))); }
void put_fixed64(std::string& s, uint64_t v) { for (int i = 0; i < 8; ++i) s.push_back((char)(v >> (8 * i))); }
bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t& v)
{
    v = 0;
    for (int shift = 0; shift <= 63 && p < end; shift += 7) {
        uint64_t b = *p++;
        v |= (b & 127) << shift;
        if (!(b & 128)) return true;
    }
    return false;
}
uint32_t get_fixed32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// ---- the two protobuf messages, by field number (tensor_bundle.proto, tensor_shape.proto)
struct Entry {
    int dtype = 0;                     // DataType: DT_FLOAT = 1
    std::vector<int64_t> shape;
    int shard = 0;
    int64_t offset = 0, size = 0;
    uint32_t crc = 0;                  // masked CRC32C of the tensor bytes
    bool sliced = false;
};

bool skip_field(const uint8_t*& p, const uint8_t* end, int wire)
{
    uint64_t v;
    switch (wire) {
    case 0: return get_varint(p, end, v);
    case 1: if (end - p < 8) return false; p += 8; return true;
    case 2: if (!get_varint(p, end, v) || (uint64_t)(end - p) < v) return false; p += v; return true;
    case 5: if (end - p < 4) return false; p += 4; return true;
    default: return false;
    }
}

bool parse_shape(const uint8_t* p, const uint8_t* end, std::vector<int64_t>& shape)
{
    while (p < end) {
        uint64_t tag; if (!get_varint(p, end, tag)) return false;
        if (tag == ((2u << 3) | 2u)) {                       // repeated Dim dim = 2
            uint64_t len; if (!get_varint(p, end, len) || (uint64_t)(end - p) < len) return false;
            const uint8_t* q = p; const uint8_t* qe = p + len; p = qe;
            int64_t size = 0;
            while (q < qe) {
                uint64_t t2; if (!get_varint(q, qe, t2)) return false;
                if (t2 == ((1u << 3) | 0u)) { uint64_t v; if (!get_varint(q, qe, v)) return false; size = (int64_t)v; }
                else if (!skip_field(q, qe, (int)(t2 & 7))) return false;
            }
            shape.push_back(size);
        } else if (!skip_field(p, end, (int)(tag & 7))) return false;
    }
    return true;
}

bool parse_entry(const std::string& s, Entry& e)
{
    const uint8_t* p = (const uint8_t*)s.data(); const uint8_t* end = p + s.size();
    while (p < end) {
        uint64_t tag, v; if (!get_varint(p, end, tag)) return false;
        const int field = (int)(tag >> 3), wire = (int)(tag & 7);
        if (field == 1 && wire == 0) { if (!get_varint(p, end, v)) return false; e.dtype = (int)v; }
        else if (field == 2 && wire == 2) {
            if (!get_varint(p, end, v) || (uint64_t)(end - p) < v) return false;
            if (!parse_shape(p, p + v, e.shape)) return false;
            p += v;
        }
        else if (field == 3 && wire == 0) { if (!get_varint(p, end, v)) return false; e.shard = (int)v; }
        else if (field == 4 && wire == 0) { if (!get_varint(p, end, v)) return false; e.offset = (int64_t)v; }
        else if (field == 5 && wire == 0) { if (!get_varint(p, end, v)) return false; e.size = (int64_t)v; }
        else if (field == 6 && wire == 5) { if (end - p < 4) return false; e.crc = get_fixed32(p); p += 4; }
        else { if (field == 7) e.sliced = true; if (!skip_field(p, end, wire)) return false; }
    }
    return true;
}

std::string encode_entry(const Entry& e)
{
    std::string shape;
    for (int64_t d : e.shape) {
        std::string dim; dim.push_back((char)((1 << 3) | 0)); put_varint(dim, (uint64_t)d);
        shape.push_back((char)((2 << 3) | 2)); put_varint(shape, dim.size()); shape += dim;
    }
    std::string s;
    s.push_back((char)((1 << 3) | 0)); put_varint(s, (uint64_t)e.dtype);
    s.push_back((char)((2 << 3) | 2)); put_varint(s, shape.size()); s += shape;
    if (e.shard) { s.push_back((char)((3 << 3) | 0)); put_varint(s, (uint64_t)e.shard); }
    if (e.offset) { s.push_back((char)((4 << 3) | 0)); put_varint(s, (uint64_t)e.offset); }
    s.push_back((char)((5 << 3) | 0)); put_varint(s, (uint64_t)e.size);
    s.push_back((char)((6 << 3) | 5)); put_fixed32(s, e.crc);
    return s;
}

// BundleHeaderProto { num_shards = 1; endianness = LITTLE (0, default: not written); version { producer = 1 } }
std::string encode_header()
{
    std::string s;
    s.push_back((char)((1 << 3) | 0)); put_varint(s, 1);
    std::string ver; ver.push_back((char)((1 << 3) | 0)); put_varint(ver, 1);
    s.push_back((char)((3 << 3) | 2)); put_varint(s, ver.size()); s += ver;
    return s;
}

bool parse_header(const std::string& s, int& num_shards, int& endianness)
{
    const uint8_t* p = (const uint8_t*)s.data(); const uint8_t* end = p + s.size();
    num_shards = 0; endianness = 0;
    while (p < end) {
        uint64_t tag, v; if (!get_varint(p, end, tag)) return false;
        if (tag == ((1u << 3) | 0u)) { if (!get_varint(p, end, v)) return false; num_shards = (int)v; }
        else if (tag == ((2u << 3) | 0u)) { if (!get_varint(p, end, v)) return false; endianness = (int)v; }
        else if (!skip_field(p, end, (int)(tag & 7))) return false;
    }
    return true;
}

// ---- table blocks
struct BlockBuilder {
    std::string buf, last_key;
    std::vector<uint32_t> restarts{ 0 };
    int counter = 0;
    bool empty() const { return buf.empty(); }
    size_t size_estimate() const { return buf.size() + restarts.size() * 4 + 4; }
    void add(const std::string& key, const std::string& value)
    {
        size_t shared = 0;
        if (counter < kRestartInterval) {
            const size_t m = key.size() < last_key.size() ? key.size() : last_key.size();
            while (shared < m && key[shared] == last_key[shared]) ++shared;
        } else { restarts.push_back((uint32_t)buf.size()); counter = 0; }
        put_varint(buf, shared); put_varint(buf, key.size() - shared); put_varint(buf, value.size());
        buf.append(key, shared, std::string::npos); buf += value;
        last_key = key; ++counter;
    }
    std::string finish()
    {
        std::string out = buf;
        for (uint32_t r : restarts) put_fixed32(out, r);
        put_fixed32(out, (uint32_t)restarts.size());
        return out;
    }
};

// appends block + trailer to `file`, returns the BlockHandle encoding
std::string write_block(std::string& file, const std::string& contents)
{
    std::string handle; put_varint(handle, file.size()); put_varint(handle, contents.size());
    file += contents;
    const char type = 0;                                     // kNoCompression
    uint32_t crc = crc32c_extend(crc32c_extend(0, contents.data(), contents.size()), &type, 1);
    file.push_back(type); put_fixed32(file, crc_mask(crc));
    return handle;
}

bool read_file(const std::string& path, std::string& out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    bool ok = n >= 0 && (n == 0 || fread(&out[0], 1, (size_t)n, f) == (size_t)n);
    fclose(f);
    return ok;
}

bool write_file(const std::string& path, const std::string& data)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = data.empty() || fwrite(data.data(), 1, data.size(), f) == data.size();
    return (fclose(f) == 0) && ok;
}

// one block of `file` at handle (offset, size): checks the trailer, walks the entries
bool read_block(const std::string& file, uint64_t off, uint64_t size, std::vector<std::pair<std::string, std::string>>& out, std::string& why)
{
    // no additions on file-supplied values (a hostile handle such as off = 2^64 - 10 must not wrap past the test)
    if (size < 4 || file.size() < 5 || size > file.size() - 5 || off > file.size() - 5 - size) { why = "block handle outside the file"; return false; }
    const uint8_t* b = (const uint8_t*)file.data() + off;
    if (b[size] != 0) { why = "compressed table block (snappy) is not supported"; return false; }
    if (crc_unmask(get_fixed32(b + size + 1)) != crc32c_extend(0, b, size + 1)) { why = "table block checksum mismatch"; return false; }
    const uint32_t n_restarts = get_fixed32(b + size - 4);
    if ((uint64_t)n_restarts * 4 + 4 > size) { why = "bad restart array"; return false; }
    const uint8_t* p = b; const uint8_t* end = b + size - 4 - (size_t)n_restarts * 4;
    std::string key;
    while (p < end) {
        uint64_t shared, unshared, vlen;
        if (!get_varint(p, end, shared) || !get_varint(p, end, unshared) || !get_varint(p, end, vlen) || shared > key.size() ||
            (uint64_t)(end - p) < unshared + vlen) { why = "corrupt table entry"; return false; }
        key.resize(shared); key.append((const char*)p, unshared); p += unshared;
        out.emplace_back(key, std::string((const char*)p, vlen)); p += vlen;
    }
    return true;
}

}  // namespace

struct az_ckpt {
    std::string prefix;
    std::vector<std::string> names;            // ascending (table order)
    std::map<std::string, Entry> entries;
    int num_shards = 1;
    std::vector<std::string> shard_cache;      // data shards, loaded on first read
};

static std::string shard_path(const std::string& prefix, int shard, int num_shards)
{
    char buf[64]; snprintf(buf, sizeof buf, ".data-%05d-of-%05d", shard, num_shards);
    return prefix + buf;
}

extern "C" uint32_t az_crc32c(const void* data, size_t n) { return crc32c_extend(0, data, n); }

extern "C" int az_ckpt_open(const char* prefix, az_ckpt** out)
{
    if (!prefix || !out) { az_set_error("NULL argument"); return AZ_ERR_INVALID_ARG; }
    std::string file;
    if (!read_file(std::string(prefix) + ".index", file)) { az_set_error("cannot read %s.index", prefix); return AZ_ERR_INVALID_ARG; }
    if (file.size() < 48) { az_set_error("%s.index is shorter than a table footer", prefix); return AZ_ERR_INVALID_ARG; }
    const uint8_t* foot = (const uint8_t*)file.data() + file.size() - 48;
    uint64_t magic = 0; for (int i = 7; i >= 0; --i) magic = (magic << 8) | foot[40 + i];
    if (magic != kTableMagic) { az_set_error("%s.index: bad table magic", prefix); return AZ_ERR_INVALID_ARG; }
    const uint8_t* p = foot; const uint8_t* pe = foot + 40;
    uint64_t mo, ms, io, is;
    if (!get_varint(p, pe, mo) || !get_varint(p, pe, ms) || !get_varint(p, pe, io) || !get_varint(p, pe, is)) {
        az_set_error("%s.index: bad footer", prefix); return AZ_ERR_INVALID_ARG; }
    std::string why;
    std::vector<std::pair<std::string, std::string>> index, kv;
    if (!read_block(file, io, is, index, why)) { az_set_error("%s.index: %s (index block)", prefix, why.c_str()); return AZ_ERR_INVALID_ARG; }
    for (auto& ie : index) {
        const uint8_t* h = (const uint8_t*)ie.second.data(); const uint8_t* he = h + ie.second.size();
        uint64_t bo, bs;
        if (!get_varint(h, he, bo) || !get_varint(h, he, bs)) { az_set_error("%s.index: bad block handle", prefix); return AZ_ERR_INVALID_ARG; }
        if (!read_block(file, bo, bs, kv, why)) { az_set_error("%s.index: %s", prefix, why.c_str()); return AZ_ERR_INVALID_ARG; }
    }
    if (kv.empty() || !kv[0].first.empty()) { az_set_error("%s.index: no bundle header entry", prefix); return AZ_ERR_INVALID_ARG; }
    az_ckpt* c = new (std::nothrow) az_ckpt();
    if (!c) { az_set_error("out of host memory"); return AZ_ERR_INVALID_ARG; }
    c->prefix = prefix;
    int endian = 0;
    if (!parse_header(kv[0].second, c->num_shards, endian) || c->num_shards < 1 || endian != 0) {
        az_set_error("%s.index: unsupported bundle header (shards %d, endianness %d)", prefix, c->num_shards, endian); delete c; return AZ_ERR_INVALID_ARG; }
    for (size_t i = 1; i < kv.size(); ++i) {
        Entry e;
        if (!parse_entry(kv[i].second, e) || e.shard < 0 || e.shard >= c->num_shards) {
            az_set_error("%s.index: bad entry for '%s'", prefix, kv[i].first.c_str()); delete c; return AZ_ERR_INVALID_ARG; }
        c->names.push_back(kv[i].first);
        c->entries[kv[i].first] = e;
    }
    c->shard_cache.resize((size_t)c->num_shards);
    *out = c;
    return AZ_OK;
}

extern "C" int az_ckpt_close(az_ckpt* c) { delete c; return AZ_OK; }
extern "C" int az_ckpt_num_tensors(const az_ckpt* c) { return c ? (int)c->names.size() : 0; }

extern "C" int az_ckpt_tensor_info(const az_ckpt* c, int i, const char** name, int* dtype, int* rank, int64_t* shape8, size_t* bytes)
{
    if (!c || i < 0 || i >= (int)c->names.size()) { az_set_error("tensor index out of range"); return AZ_ERR_INVALID_ARG; }
    const Entry& e = c->entries.at(c->names[(size_t)i]);
    if (name) *name = c->names[(size_t)i].c_str();
    if (dtype) *dtype = e.dtype;
    if (rank) *rank = (int)e.shape.size();
    if (shape8) for (size_t k = 0; k < 8; ++k) shape8[k] = k < e.shape.size() ? e.shape[k] : 0;
    if (bytes) *bytes = (size_t)e.size;
    return AZ_OK;
}

extern "C" int az_ckpt_find(const az_ckpt* c, const char* name)
{
    if (!c || !name) return -1;
    for (size_t i = 0; i < c->names.size(); ++i) if (c->names[i] == name) return (int)i;
    return -1;
}

extern "C" int az_ckpt_read(az_ckpt* c, const char* name, void* h_out, size_t bytes)
{
    if (!c || !name || !h_out) { az_set_error("NULL argument"); return AZ_ERR_INVALID_ARG; }
    auto it = c->entries.find(name);
    if (it == c->entries.end()) { az_set_error("checkpoint %s has no tensor '%s'", c->prefix.c_str(), name); return AZ_ERR_INVALID_ARG; }
    const Entry& e = it->second;
    if (e.sliced) { az_set_error("tensor '%s' is stored in slices (partitioned variable): not supported", name); return AZ_ERR_INVALID_ARG; }
    if ((size_t)e.size != bytes) { az_set_error("tensor '%s' holds %lld bytes, caller asked for %zu", name, (long long)e.size, bytes); return AZ_ERR_INVALID_ARG; }
    std::string& shard = c->shard_cache[(size_t)e.shard];
    if (shard.empty() && !read_file(shard_path(c->prefix, e.shard, c->num_shards), shard)) {
        az_set_error("cannot read %s", shard_path(c->prefix, e.shard, c->num_shards).c_str()); return AZ_ERR_INVALID_ARG; }
    if (e.offset < 0 || e.size < 0 || (uint64_t)e.size > shard.size() || (uint64_t)e.offset > shard.size() - (uint64_t)e.size) {
        az_set_error("tensor '%s' lies outside its data shard", name); return AZ_ERR_INVALID_ARG; }
    if (crc_unmask(e.crc) != crc32c_extend(0, shard.data() + e.offset, (size_t)e.size)) {
        az_set_error("tensor '%s': data checksum mismatch", name); return AZ_ERR_BAD_STATE; }
    memcpy(h_out, shard.data() + e.offset, bytes);
    return AZ_OK;
}

// BundleWriter: float32 tensors, one shard, entries and data in ascending name order (what the Saver's sorted save op produces)
extern "C" int az_ckpt_write(const char* prefix, int n, const char* const* names, const int* ranks, const int64_t* const* shapes,
                             const float* const* data)
{
    if (!prefix || n < 0 || (n > 0 && (!names || !ranks || !shapes || !data))) { az_set_error("NULL argument"); return AZ_ERR_INVALID_ARG; }
    std::map<std::string, int> order;
    for (int i = 0; i < n; ++i) {
        if (!names[i] || !names[i][0] || ranks[i] < 0 || ranks[i] > 8) { az_set_error("bad tensor %d", i); return AZ_ERR_INVALID_ARG; }
        if (!order.emplace(names[i], i).second) { az_set_error("duplicate tensor name '%s'", names[i]); return AZ_ERR_INVALID_ARG; }
    }
    std::string blob, table;
    BlockBuilder block, index;
    auto flush = [&]() {
        if (block.empty()) return;
        const std::string last = block.last_key;
        index.add(last, write_block(table, block.finish()));            // the block's last key is a valid separator
        block = BlockBuilder();
    };
    block.add("", encode_header());
    for (auto& kv : order) {
        const int i = kv.second;
        Entry e; e.dtype = 1;
        size_t count = 1;
        for (int k = 0; k < ranks[i]; ++k) { e.shape.push_back(shapes[i][k]); count *= (size_t)shapes[i][k]; }
        e.offset = (int64_t)blob.size(); e.size = (int64_t)(count * sizeof(float));
        e.crc = crc_mask(crc32c_extend(0, data[i], count * sizeof(float)));
        blob.append((const char*)data[i], count * sizeof(float));
        block.add(kv.first, encode_entry(e));
        if (block.size_estimate() >= kBlockSize) flush();
    }
    flush();
    BlockBuilder meta;
    std::string foot = write_block(table, meta.finish());
    foot += write_block(table, index.finish());
    foot.resize(40, '\0');
    put_fixed64(foot, kTableMagic);
    table += foot;
    // TensorFlow writes the shard first, the index last (a reader that sees the index may rely on the data)
    if (!write_file(shard_path(prefix, 0, 1), blob) || !write_file(std::string(prefix) + ".index", table)) {
        az_set_error("cannot write checkpoint %s", prefix); return AZ_ERR_INVALID_ARG; }
    return AZ_OK;
}
