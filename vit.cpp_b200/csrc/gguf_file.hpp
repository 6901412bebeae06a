// gguf_file.hpp -- reader for a true GGUF (v2/v3) container holding a ViT (host code, no CUDA).
//
// The reference's ".gguf" files are the legacy ggml format (magic "ggml", reference vit.cpp:320-328); real GGUF exists in its ggml
// tree (ggml/docs/gguf.md, gguf_init_from_file ggml.c:18017) but vit.cpp never calls it and defines no ViT key names.  This
// reader is the SURVEY.md 8(f) rank-3 extension: same tensors (timm names, ggml dimension order ne[0] = fastest), same types
// (+ BF16, ggml type 30 in current ggml, which gives bf16 checkpoints a container), hyper-parameters as metadata:
//     general.architecture = "vit"                         general.alignment (default 32)          general.file_type (optional)
//     vit.hidden_size, vit.num_hidden_layers, vit.num_attention_heads, vit.num_classes, vit.patch_size, vit.image_size  (integers)
//     vit.layer_norm_eps (float32, default 1e-6)           vit.id2label (array of strings, index = class id; optional)
// Layout (gguf.md): header {magic "GGUF", u32 version, u64 n_tensors, u64 n_kv}, n_kv x {string key, u32 type, value},
// n_tensors x {string name, u32 n_dims, u64 ne[n_dims], u32 type, u64 offset}, padding to the alignment, tensor data (offsets
// relative to the start of the data section).  Strings are {u64 length, bytes}.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace vitb200 {

struct GgufTensor
{
    std::string name;
    int32_t type = 0, n_dims = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    size_t offset = 0, nbytes = 0; // offset from the start of the file
};

struct GgufModel
{
    int64_t hidden_size = 0, num_hidden_layers = 0, num_attention_heads = 0, num_classes = 0, patch_size = 0, img_size = 0, ftype = 1;
    float eps = 1e-6f;
    std::map<int, std::string> labels;
    std::vector<GgufTensor> tensors;
    std::string error;
};

constexpr uint32_t GGUF_MAGIC_LE = 0x46554747u; // "GGUF"

// bytes of n elements of a ggml tensor type; 0 = unsupported type or a row that is not a whole number of blocks
inline size_t gguf_type_bytes(int type, int64_t n, int64_t ne0)
{
    switch (type)
    {
    case 0: return (size_t)n * 4;   // F32
    case 1: return (size_t)n * 2;   // F16
    case 30: return (size_t)n * 2;  // BF16
    case 2: return ne0 % 32 ? 0 : (size_t)n / 32 * 18; // Q4_0
    case 3: return ne0 % 32 ? 0 : (size_t)n / 32 * 20; // Q4_1
    case 6: return ne0 % 32 ? 0 : (size_t)n / 32 * 22; // Q5_0
    case 7: return ne0 % 32 ? 0 : (size_t)n / 32 * 24; // Q5_1
    case 8: return ne0 % 32 ? 0 : (size_t)n / 32 * 34; // Q8_0
    default: return 0;
    }
}

inline bool parse_gguf(const char *buf, size_t size, GgufModel &m)
{
    size_t off = 0;
    auto fail = [&](const std::string &why) { m.error = why; return false; };
    auto rd = [&](void *dst, size_t n) { if (off + n > size) return false; memcpy(dst, buf + off, n); off += n; return true; };
    auto rd_str = [&](std::string &s) {
        uint64_t len = 0;
        if (!rd(&len, 8) || len > size - off) return false;
        s.assign(buf + off, (size_t)len);
        off += (size_t)len;
        return true;
    };
    uint32_t magic = 0, version = 0;
    uint64_t n_tensors = 0, n_kv = 0;
    if (!rd(&magic, 4) || magic != GGUF_MAGIC_LE) return fail("not a GGUF file");
    if (!rd(&version, 4) || (version != 2 && version != 3)) return fail("unsupported GGUF version " + std::to_string(version));
    if (!rd(&n_tensors, 8) || !rd(&n_kv, 8) || n_tensors > (1u << 20) || n_kv > (1u << 20)) return fail("corrupt GGUF header");

    static const size_t scalar_size[13] = {1, 1, 2, 2, 4, 4, 4, 1, 0, 0, 8, 8, 8};
    uint64_t alignment = 32;
    std::string arch;
    for (uint64_t i = 0; i < n_kv; ++i)
    {
        std::string key;
        uint32_t vt = 0;
        if (!rd_str(key) || !rd(&vt, 4) || vt > 12) return fail("corrupt GGUF metadata");
        if (vt == 8) // string
        {
            std::string v;
            if (!rd_str(v)) return fail("truncated GGUF string value");
            if (key == "general.architecture") arch = v;
        }
        else if (vt == 9) // array
        {
            uint32_t et = 0;
            uint64_t cnt = 0;
            if (!rd(&et, 4) || !rd(&cnt, 8) || et > 12 || et == 9) return fail("corrupt GGUF array");
            for (uint64_t j = 0; j < cnt; ++j)
            {
                if (et == 8)
                {
                    std::string v;
                    if (!rd_str(v)) return fail("truncated GGUF string array");
                    if (key == "vit.id2label") m.labels[(int)j] = v;
                }
                else
                {
                    if (off + scalar_size[et] > size) return fail("truncated GGUF array");
                    off += scalar_size[et];
                }
            }
        }
        else
        {
            unsigned char raw[8] = {0};
            if (!rd(raw, scalar_size[vt])) return fail("truncated GGUF value");
            int64_t iv = 0;
            double fv = 0.0;
            switch (vt)
            {
            case 0: iv = *(uint8_t *)raw; break;
            case 1: iv = *(int8_t *)raw; break;
            case 2: { uint16_t x; memcpy(&x, raw, 2); iv = x; break; }
            case 3: { int16_t x; memcpy(&x, raw, 2); iv = x; break; }
            case 4: { uint32_t x; memcpy(&x, raw, 4); iv = x; break; }
            case 5: { int32_t x; memcpy(&x, raw, 4); iv = x; break; }
            case 6: { float x; memcpy(&x, raw, 4); fv = x; iv = (int64_t)x; break; }
            case 7: iv = raw[0] != 0; break;
            case 10: { uint64_t x; memcpy(&x, raw, 8); iv = (int64_t)x; break; }
            case 11: memcpy(&iv, raw, 8); break;
            case 12: memcpy(&fv, raw, 8); iv = (int64_t)fv; break;
            }
            if (vt != 6 && vt != 12) fv = (double)iv;
            if (key == "general.alignment") alignment = (uint64_t)iv;
            else if (key == "general.file_type") m.ftype = iv;
            else if (key == "vit.hidden_size") m.hidden_size = iv;
            else if (key == "vit.num_hidden_layers" || key == "vit.block_count") m.num_hidden_layers = iv;
            else if (key == "vit.num_attention_heads") m.num_attention_heads = iv;
            else if (key == "vit.num_classes") m.num_classes = iv;
            else if (key == "vit.patch_size") m.patch_size = iv;
            else if (key == "vit.image_size" || key == "vit.img_size") m.img_size = iv;
            else if (key == "vit.layer_norm_eps") m.eps = (float)fv;
        }
    }
    if (!arch.empty() && arch != "vit") return fail("GGUF architecture is '" + arch + "', expected 'vit'");
    if (alignment < 8 || alignment % 8 != 0 || alignment > (1u << 20)) return fail("bad general.alignment");
    if (m.hidden_size <= 0 || m.num_hidden_layers <= 0 || m.num_attention_heads <= 0 || m.num_classes <= 0 || m.patch_size <= 0 || m.img_size <= 0)
        return fail("GGUF file lacks the vit.* hyper-parameter keys");

    m.tensors.resize((size_t)n_tensors);
    for (auto &t : m.tensors)
    {
        uint32_t nd = 0, type = 0;
        uint64_t rel = 0;
        if (!rd_str(t.name) || !rd(&nd, 4) || nd < 1 || nd > 4) return fail("corrupt GGUF tensor info");
        int64_t n = 1;
        for (uint32_t d = 0; d < nd; ++d)
        {
            uint64_t e = 0;
            if (!rd(&e, 8) || e < 1 || e > (1ull << 40)) return fail("corrupt GGUF tensor dims");
            t.ne[d] = (int64_t)e;
            if (n > (int64_t)(1ull << 40) / (int64_t)e) return fail("corrupt GGUF tensor dims (too many elements)");
            n *= (int64_t)e;
        }
        if (!rd(&type, 4) || !rd(&rel, 8)) return fail("corrupt GGUF tensor info");
        t.n_dims = (int32_t)nd;
        t.type = (int32_t)type;
        t.nbytes = gguf_type_bytes(t.type, n, t.ne[0]);
        if (t.nbytes == 0) return fail("tensor '" + t.name + "': unsupported ggml type " + std::to_string(type));
        if (rel % alignment != 0) return fail("tensor '" + t.name + "': data offset is not aligned");
        t.offset = (size_t)rel; // made absolute below
    }
    const size_t data0 = off + (alignment - off % alignment) % alignment;
    for (auto &t : m.tensors)
    {
        t.offset += data0;
        if (t.offset + t.nbytes > size) return fail("tensor '" + t.name + "' has wrong size in model file");
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Legacy-ggml model file (the format reference vit_model_load parses, vit.cpp:308-712; written by convert-pth-to-ggml.py:105-158):
//   int32 magic 0x67676d6c ("ggml", ggml.h:211), 7 x int32 hparams (hidden, layers, heads, classes, patch, img, ftype; vit.cpp:335-341),
//   int32 n_labels, n x {int32 id, int32 len, bytes} (vit.cpp:356-371), then per tensor {int32 n_dims, int32 name_len, int32 type,
//   n_dims x int32 ne, name, raw data} (vit.cpp:590-695).  Validation mirrors the reference (magic, known types, sizes); the tensor
//   inventory and shapes are checked against the hyper-parameters by vitb200_create.  Fills the same GgufModel.
inline bool parse_legacy_ggml(const char *buf, size_t size, GgufModel &m)
{
    size_t off = 0;
    auto fail = [&](const std::string &why) { m.error = why; return false; };
    auto rd32 = [&](int32_t &v) { if (off + 4 > size) return false; memcpy(&v, buf + off, 4); off += 4; return true; };
    int32_t magic = 0;
    if (!rd32(magic) || (uint32_t)magic != 0x67676d6cu) return fail("bad magic");
    int32_t hp[7] = {0};
    for (int i = 0; i < 7; ++i)
        if (!rd32(hp[i])) return fail("truncated header");
    m.hidden_size = hp[0]; m.num_hidden_layers = hp[1]; m.num_attention_heads = hp[2]; m.num_classes = hp[3];
    m.patch_size = hp[4]; m.img_size = hp[5];
    m.ftype = hp[6] % 1000; // GGML_QNT_VERSION_FACTOR, vit.cpp:343-354
    m.eps = 1e-6f;
    int32_t n_labels = 0;
    if (!rd32(n_labels) || n_labels < 0) return fail("truncated label table");
    for (int i = 0; i < n_labels; ++i)
    {
        int32_t key = 0, len = 0;
        if (!rd32(key) || !rd32(len) || len < 0 || (size_t)len > size - off) return fail("truncated label table");
        m.labels[key] = std::string(buf + off, (size_t)len);
        off += (size_t)len;
    }
    while (off < size)
    {
        GgufTensor t;
        int32_t len = 0;
        if (!rd32(t.n_dims) || !rd32(len) || !rd32(t.type)) return fail("truncated tensor record");
        if (t.n_dims < 1 || t.n_dims > 4 || len < 0) return fail("corrupt tensor record");
        int64_t n = 1;
        for (int i = 0; i < t.n_dims; ++i)
        {
            int32_t d = 0;
            if (!rd32(d) || d < 1) return fail("corrupt tensor dims");
            t.ne[i] = d;
            if (n > (int64_t)(1ull << 40) / d) return fail("corrupt tensor dims (too many elements)");
            n *= d;
        }
        if ((size_t)len > size - off) return fail("truncated tensor name");
        t.name.assign(buf + off, (size_t)len);
        off += (size_t)len;
        if (t.type == 30) return fail("unknown ftype 30 in model file (tensor '" + t.name + "')"); // BF16 exists in GGUF containers only
        t.nbytes = gguf_type_bytes(t.type, n, t.ne[0]); // vit.cpp:645-678
        if (t.nbytes == 0)
            return fail(t.type == 0 || t.type == 1 || t.type == 2 || t.type == 3 || t.type == 6 || t.type == 7 || t.type == 8
                            ? "tensor '" + t.name + "': quantised rows must be a multiple of 32"
                            : "unknown ftype " + std::to_string(t.type) + " in model file (tensor '" + t.name + "')");
        if (t.nbytes > size - off) return fail("tensor '" + t.name + "' has wrong size in model file");
        t.offset = off;
        off += t.nbytes;
        m.tensors.push_back(t);
    }
    return true;
}

} // namespace vitb200
