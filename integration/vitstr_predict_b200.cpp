// integration/vitstr_predict_b200.cpp -- the reference-side binding for the ViTSTR extension (extensions/vitstr.cpp).
//
// Compiled against the extension's own vitstr.h, this translation unit provides vit_predict (vitstr.h:123,
// vitstr.cpp:979-1061) and a vit_encode_image stub on top of the C ABI in include/vitb200.h.  The extension's loader,
// grayscale preprocess, argv parsing and main.cpp stay its own code: build its vitstr.cpp with
//     -Dvit_predict=vitstr_predict_ggml_ref -Dvit_encode_image=vitstr_encode_image_ggml_ref
// and link this file + libvitb200.so in their place (oracle/Makefile target `cli`).
#include "vitstr.h"

#include "engine_cache.hpp"

#include <cstdio>
#include <vector>

namespace {

const int kSeqLen = 25; // tokens the classifier reads (vitstr.cpp:865)
vitb200_shim::EngineCache<vit_model> g_engines; // the [P, P, 1, D] patch kernel (vitstr.cpp:482) tells the engine the input is 1-channel

vitb200_engine *engine_for(const vit_model &model)
{
    return g_engines.get(model, vitb200_shim::env_int("VITB200_DEVICE", 0), vitb200_shim::env_int("VITB200_MAX_BATCH", 16), kSeqLen);
}

} // namespace

void vit_b200_release(const vit_model *model) { g_engines.release(model); }

struct ggml_cgraph *vit_encode_image(const vit_model &, vit_state &, const image_f32 &) { return nullptr; }

// Drop-in vit_predict of the extension: state.prediction ([num_classes, 25]) is filled with the per-token probabilities and
// the greedy decode is printed the way the extension prints it (stop at class 1 = end of sentence, score = product of the
// per-character maxima); params.n_threads is ignored on the GPU path.
int vit_predict(const vit_model &model, vit_state &state, const image_f32 img1, const vit_params &params,
                std::vector<std::pair<float, int>> &predictions)
{
    (void)params;
    vitb200_engine *e = engine_for(model);
    if (!e)
    {
        fprintf(stderr, "%s: failed to create the GPU engine: %s\n", __func__, vitb200_last_error());
        return 1;
    }
    const int nc = model.hparams.num_classes, S = model.hparams.img_size;
    if (img1.nx != S || img1.ny != S || img1.data.size() < (size_t)S * S)
    {
        fprintf(stderr, "%s: the image must be %d x %d grayscale\n", __func__, S, S);
        return 1;
    }
    std::vector<float> local;
    float *probs = (state.prediction && state.prediction->data) ? ggml_get_data_f32(state.prediction) : nullptr;
    if (!probs) { local.resize((size_t)kSeqLen * nc); probs = local.data(); }
    if (vitb200_forward(e, img1.data.data(), 1, probs, nullptr, nullptr, nullptr, 0) != 0)
    {
        fprintf(stderr, "%s: %s\n", __func__, vitb200_last_error());
        return 1;
    }
    predictions.clear();
    printf("------------------ \n");
    double score = 1.0;
    for (int t = 1; t < kSeqLen; ++t) // token 0 is the [GO] position
    {
        const float *row = probs + (size_t)t * nc;
        int best = 0;
        for (int c = 1; c < nc; ++c)
            if (row[c] > row[best]) best = c;
        if (best == 1) break; // [s]
        score *= row[best];
        auto it = model.hparams.id2label.find(best);
        printf("%s", it != model.hparams.id2label.end() ? it->second.c_str() : "?");
    }
    printf("\n");
    printf("score : %.2f \n", score);
    printf("------------------ \n");
    return 0;
}
