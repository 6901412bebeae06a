// integration/engine_cache.hpp -- device-copy cache shared by the two reference-side bindings (vit_predict_b200.cpp,
// vitstr_predict_b200.cpp).  Both reference headers declare their own `vit_model` (vit.h:82-89, vitstr.h), hence the template.
//
// Device copies are cached per model (SURVEY.md 8b "Ownership").  The key is the vit_model's address, but an address alone is not
// an identity: a caller may free a model (ggml_free(model.ctx), main.cpp:110) and load another one into the same object.  Every
// entry therefore carries a fingerprint of the host weights it was built from -- the owning ggml context, the hyper-parameters, and
// for every tensor its data pointer plus a hash of its first and last bytes -- and is rebuilt when the fingerprint no longer
// matches.  release() frees one device copy explicitly (call it next to ggml_free(model.ctx)); what is left is freed at exit.
#pragma once
#include "vitb200.h"

#include <cstdint>
#include <cstdlib>
#include <map>
#include <vector>

namespace vitb200_shim {

inline uint64_t fnv1a(uint64_t h, const void *p, size_t n)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

template <typename Model>
uint64_t model_fingerprint(const Model &model)
{
    uint64_t h = 1469598103934665603ull;
    const void *ctx = model.ctx;
    h = fnv1a(h, &ctx, sizeof(ctx));
    const int32_t hp[7] = {model.hparams.hidden_size, model.hparams.num_hidden_layers, model.hparams.num_attention_heads, model.hparams.num_classes,
                           model.hparams.patch_size, model.hparams.img_size, model.hparams.ftype};
    h = fnv1a(h, hp, sizeof(hp));
    for (const auto &kv : model.tensors)
    {
        const void *d = kv.second->data;
        h = fnv1a(h, kv.first.data(), kv.first.size());
        h = fnv1a(h, &d, sizeof(d));
        const size_t nb = ggml_nbytes(kv.second);
        if (d) h = fnv1a(h, d, nb < 256 ? nb : 256);
        if (d && nb > 512) h = fnv1a(h, (const char *)d + nb - 256, 256);
    }
    return h;
}

template <typename Model>
class EngineCache
{
    struct Entry { vitb200_engine *engine; uint64_t fingerprint; };
    std::map<const Model *, Entry> entries_;

  public:
    ~EngineCache() { release(nullptr); } // static storage duration: runs at exit

    void release(const Model *model) // nullptr: every model
    {
        if (!model)
        {
            for (auto &kv : entries_) vitb200_destroy(kv.second.engine);
            entries_.clear();
            return;
        }
        auto it = entries_.find(model);
        if (it == entries_.end()) return;
        vitb200_destroy(it->second.engine);
        entries_.erase(it);
    }

    // the engine holding `model`'s weights, created (vitb200_create_ex with what vit_model::tensors holds, vit.h:88) on first use
    vitb200_engine *get(const Model &model, int device, int max_batch, int head_tokens)
    {
        const uint64_t fp = model_fingerprint(model);
        auto it = entries_.find(&model);
        if (it != entries_.end())
        {
            if (it->second.fingerprint == fp) return it->second.engine;
            vitb200_destroy(it->second.engine); // the object now holds different weights: the device copy is stale
            entries_.erase(it);
        }
        vitb200_hparams hp;
        hp.hidden_size = model.hparams.hidden_size;
        hp.num_hidden_layers = model.hparams.num_hidden_layers;
        hp.num_attention_heads = model.hparams.num_attention_heads;
        hp.num_classes = model.hparams.num_classes;
        hp.patch_size = model.hparams.patch_size;
        hp.img_size = model.hparams.img_size;
        hp.ftype = model.hparams.ftype;
        hp.eps = model.hparams.eps;
        std::vector<vitb200_tensor> ts;
        for (const auto &kv : model.tensors) // names at vit.cpp:518-579; the [P, P, C, D] patch kernel tells the engine the channel count
        {
            vitb200_tensor t;
            t.name = kv.first.c_str();
            t.data = kv.second->data;
            t.type = (int32_t)kv.second->type; // ggml type ids: F32 = 0, F16 = 1, Q4_0 = 2, ... Q8_0 = 8
            t.n_dims = kv.second->n_dims;
            for (int i = 0; i < 4; ++i) t.ne[i] = kv.second->ne[i];
            ts.push_back(t);
        }
        vitb200_engine *e = nullptr;
        if (vitb200_create_ex(&hp, ts.data(), (int)ts.size(), device, max_batch, head_tokens, &e) != 0) return nullptr;
        entries_[&model] = Entry{e, fp};
        return e;
    }
};

inline int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

} // namespace vitb200_shim
