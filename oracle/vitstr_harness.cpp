// oracle/vitstr_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" harness around the reference's ViTSTR extension (extensions/vitstr.cpp/vitstr.cpp @ a4841f6), compiled where it
// lies by oracle/Makefile into oracle/_ref/libvitstrref.so (its own shared object: it redefines vit_model, vit_model_load,
// ... with the same names as the main reference).  The extension as shipped includes stb_image.h through an absolute path of
// its author's machine ("/home/said/projects/vit.cpp/ggml/examples/stb_image.h"); the Makefile rewrites that one include path
// to the in-tree header while piping the source into the compiler -- no reference source is copied or stored.
// As in ref_harness.cpp, -Dvit_predict=vitstr_predict_ggml_ref renames the extension's predict and
// -Dggml_soft_max=oracle_capture_soft_max_vitstr routes its single ggml_soft_max( call (vitstr.cpp:903) through the capture
// below, which keeps the 25 x num_classes pre-soft-max logits.
#include "vitstr.h" // the include-path-patched copy in oracle/_ref/obj/vitstr (build intermediate)

#include <cstdint>
#include <cstring>
#include <vector>

int vitstr_predict_ggml_ref(const vit_model &model, vit_state &state, const image_f32 img1, const vit_params &params,
                            std::vector<std::pair<float, int>> &predictions);

static std::vector<float> g_logits;

extern "C" struct ggml_tensor *oracle_capture_soft_max_vitstr(struct ggml_context *ctx, struct ggml_tensor *a)
{
    const int64_t n = ggml_nelements(a);
    if ((int64_t)g_logits.size() != n) g_logits.assign((size_t)n, 0.0f);
    struct ggml_tensor *lg = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, a->ne[0], a->ne[1]);
    lg->data = g_logits.data();
    return ggml_soft_max(ctx, ggml_cpy(ctx, a, lg));
}

struct vitstr_handle
{
    vit_model model;
    vit_state state;
};

extern "C" {

void *vitstrref_load(const char *path)
{
    vitstr_handle *h = new vitstr_handle();
    if (!vit_model_load(std::string(path), h->model))
    {
        delete h;
        return nullptr;
    }
    // same state setup as the extension's main.cpp:80-91
    struct ggml_init_params p = {3u * 1024 * 1024, NULL, false};
    h->state.ctx = ggml_init(p);
    h->state.prediction = ggml_new_tensor_2d(h->state.ctx, GGML_TYPE_F32, h->model.hparams.num_classes, 25);
    return h;
}

// out[0..6] = hidden, layers, heads, classes, patch, img, ftype
int vitstrref_hparams(void *hv, int32_t *out)
{
    vitstr_handle *h = (vitstr_handle *)hv;
    if (!h) return 1;
    const vit_hparams &hp = h->model.hparams;
    out[0] = hp.hidden_size; out[1] = hp.num_hidden_layers; out[2] = hp.num_attention_heads; out[3] = hp.num_classes;
    out[4] = hp.patch_size; out[5] = hp.img_size; out[6] = hp.ftype;
    return 0;
}

// One reference forward (the extension's vit_predict, vitstr.cpp:979) on one pre-processed grayscale image [S*S] f32.
// probs_out / logits_out: float32[25][num_classes] (either may be NULL).
int vitstrref_predict(void *hv, const float *img_gray, int n_threads, float *probs_out, float *logits_out)
{
    vitstr_handle *h = (vitstr_handle *)hv;
    if (!h) return 1;
    const vit_hparams &hp = h->model.hparams;
    image_f32 img;
    img.nx = hp.img_size;
    img.ny = hp.img_size;
    img.data.assign(img_gray, img_gray + (size_t)hp.img_size * hp.img_size);
    vit_params params;
    params.n_threads = n_threads;
    std::vector<std::pair<float, int>> predictions;
    const int rc = vitstr_predict_ggml_ref(h->model, h->state, img, params, predictions);
    if (rc != 0) return rc;
    const size_t n = (size_t)25 * hp.num_classes;
    if (probs_out) memcpy(probs_out, ggml_get_data_f32(h->state.prediction), sizeof(float) * n);
    if (logits_out) memcpy(logits_out, g_logits.data(), sizeof(float) * n);
    return 0;
}

void vitstrref_free(void *hv)
{
    vitstr_handle *h = (vitstr_handle *)hv;
    if (!h) return;
    if (h->state.ctx) ggml_free(h->state.ctx);
    if (h->model.ctx) ggml_free(h->model.ctx);
    delete h;
}

} // extern "C"
