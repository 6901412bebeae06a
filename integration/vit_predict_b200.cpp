// integration/vit_predict_b200.cpp -- the reference-side binding a vit.cpp maintainer would add.
//
// Compiled against the REFERENCE's own vit.h (found with -I<reference>), this translation unit provides the two
// symbols the forward path owns -- vit_predict (reference vit.h:122, vit.cpp:1004-1075) and vit_encode_image
// (vit.h:121) -- on top of the C ABI in include/vitb200.h.  Everything else (vit_model_load, vit_image_preprocess,
// load_image_from_file, vit_params_parse, main.cpp) stays the reference's code: build the reference's vit.cpp with
//     -Dvit_predict=vit_predict_ggml_ref -Dvit_encode_image=vit_encode_image_ggml_ref
// and link this file + libvitb200.so in their place (oracle/Makefile target `cli`).  main.cpp compiles unmodified.
#include "vit.h"

#include "engine_cache.hpp"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace {

vitb200_shim::EngineCache<vit_model> g_engines;

// new knobs come from the environment so main.cpp stays untouched (SURVEY.md section 5 "Config / flags")
vitb200_engine *engine_for(const vit_model &model)
{
    return g_engines.get(model, vitb200_shim::env_int("VITB200_DEVICE", 0), vitb200_shim::env_int("VITB200_MAX_BATCH", 256), /*head_tokens*/ 1);
}

} // namespace

// Free the device copy of one model (nullptr: of every model).  An addition next to vit_predict_batch: the reference frees a
// model with ggml_free(model.ctx) (main.cpp:110), which cannot know about device memory.
void vit_b200_release(const vit_model *model) { g_engines.release(model); }

// Kept only so the declaration in vit.h resolves; the GPU path has no ggml graph.
struct ggml_cgraph *vit_encode_image(const vit_model &, vit_state &, const image_f32 &) { return nullptr; }

// Batched entry point (an addition: the reference hard-codes one image per call, vit.cpp:747).
int vit_predict_batch(const vit_model &model, const std::vector<image_f32> &imgs, const vit_params &params,
                      std::vector<std::vector<std::pair<float, int>>> &predictions, float *probs_out)
{
    vitb200_engine *e = engine_for(model);
    if (!e)
    {
        fprintf(stderr, "%s: failed to create the GPU engine: %s\n", __func__, vitb200_last_error());
        return 1;
    }
    const int nc = model.hparams.num_classes, B = (int)imgs.size();
    const size_t per = (size_t)3 * model.hparams.img_size * model.hparams.img_size;
    std::vector<float> staging(per * B), probs((size_t)nc * B);
    for (int b = 0; b < B; ++b)
    {
        if (imgs[b].nx != model.hparams.img_size || imgs[b].ny != model.hparams.img_size || imgs[b].data.size() != per)
        {
            fprintf(stderr, "%s: image %d has the wrong size\n", __func__, b);
            return 1;
        }
        std::copy(imgs[b].data.begin(), imgs[b].data.end(), staging.begin() + per * b);
    }
    if (vitb200_forward(e, staging.data(), B, probs.data(), nullptr, nullptr, nullptr, 0) != 0)
    {
        fprintf(stderr, "%s: %s\n", __func__, vitb200_last_error());
        return 1;
    }
    predictions.assign(B, {});
    for (int b = 0; b < B; ++b)
    {
        auto &pr = predictions[b];
        pr.reserve(nc);
        for (int i = 0; i < nc; ++i) pr.push_back(std::make_pair(probs[(size_t)b * nc + i], i));
        // the reference sorts all classes descending (vit.cpp:1053-1057); stable so ties keep index order
        std::stable_sort(pr.begin(), pr.end(), [](const std::pair<float, int> &a, const std::pair<float, int> &c) { return a.first > c.first; });
    }
    if (probs_out) std::copy(probs.begin(), probs.end(), probs_out);
    (void)params;
    return 0;
}

// Drop-in vit_predict: same signature, same visible results (state.prediction filled, `predictions` full length and
// sorted, " > label : prob" lines on stdout); params.n_threads is ignored on the GPU path.
int vit_predict(const vit_model &model, vit_state &state, const image_f32 img1, const vit_params &params,
                std::vector<std::pair<float, int>> &predictions)
{
    std::vector<std::vector<std::pair<float, int>>> all;
    std::vector<image_f32> one(1, img1);
    float *dst = (state.prediction && state.prediction->data) ? ggml_get_data_f32(state.prediction) : nullptr;
    if (vit_predict_batch(model, one, params, all, dst) != 0) return 1;
    predictions = all[0];
    fprintf(stderr, "\n");
    for (int i = 0; i < params.topk && i < (int)predictions.size(); ++i)
    {
        auto it = model.hparams.id2label.find(predictions[i].second);
        printf(" > %s : %.2f\n", it != model.hparams.id2label.end() ? it->second.c_str() : "?", predictions[i].first);
    }
    return 0;
}
