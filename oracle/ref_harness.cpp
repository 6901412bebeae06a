// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" harness around the UNMODIFIED reference implementation
// (staghado/vit.cpp @ a4841f6 + vendored ggml @ a5e4560).  The reference sources are
// compiled where they lie under /root/reference by oracle/Makefile; nothing from the
// reference tree is copied into this repository.  The resulting shared object lives in
// oracle/_ref/ (git-ignored) and is used only by tests/, __graft_entry__.smoke() and the
// `cpu_baseline` / `--impl reference` legs of bench.py as the CHECKER / CPU BASELINE.
//
// How the reference's own predict becomes an in-process oracle symbol:
//   reference vit.cpp is compiled with
//     -Dvit_predict=vit_predict_ggml_ref -Dvit_encode_image=vit_encode_image_ggml_ref
//     -Dggml_soft_max=oracle_capture_soft_max
//   so (a) its loader / preprocess / CLI parsing keep their names, (b) its predict gets a
//   distinct name, and (c) the single `ggml_soft_max(` call at reference vit.cpp:931 is routed
//   through oracle_capture_soft_max() below, which copies the PRE-softmax logits into a
//   persistent host buffer before applying the real ggml_soft_max.  The attention softmax at
//   vit.cpp:856 uses a different identifier (ggml_soft_max_inplace) and is untouched.
#include "vit.h"  // reference header, found via -I/root/reference

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

// reference predict under its renamed symbol (reference vit.cpp:1004)
int vit_predict_ggml_ref(const vit_model &model, vit_state &state, const image_f32 img1,
                         const vit_params &params, std::vector<std::pair<float, int>> &predictions);

// reference graph builder under its renamed symbol (reference vit.cpp:718)
struct ggml_cgraph *vit_encode_image_ggml_ref(const vit_model &model, vit_state &state, const image_f32 &img);

static std::vector<float> g_logits;

// Called from reference vit.cpp:931 (after the -D rename).  ggml_cpy() returns a view of
// `lg`, whose data pointer we own, so ggml-alloc leaves it alone (ggml-alloc.c:469-472) and
// the logits survive the graph run.  Probabilities stay bit-identical to the stock build.
extern "C" struct ggml_tensor *oracle_capture_soft_max(struct ggml_context *ctx, struct ggml_tensor *a)
{
    const int64_t n = a->ne[0];
    if ((int64_t)g_logits.size() != n) g_logits.assign((size_t)n, 0.0f);
    struct ggml_tensor *lg = ggml_new_tensor_1d(ctx, GGML_TYPE_F32, n);
    lg->data = g_logits.data();
    return ggml_soft_max(ctx, ggml_cpy(ctx, a, lg));
}

struct vitref_handle
{
    vit_model model;
    vit_state state;
    bool loaded = false;
};

extern "C" {

// Load a legacy-ggml ViT file with the reference's own vit_model_load (vit.cpp:308).
void *vitref_load(const char *path)
{
    vitref_handle *h = new vitref_handle();
    if (!vit_model_load(std::string(path), h->model))
    {
        delete h;
        return nullptr;
    }
    // same state setup as reference main.cpp:81-94
    struct ggml_init_params p = {3u * 1024 * 1024, NULL, false};
    h->state.ctx = ggml_init(p);
    h->state.prediction = ggml_new_tensor_4d(h->state.ctx, GGML_TYPE_F32, h->model.hparams.num_classes, 1, 1, 1);
    h->loaded = true;
    return h;
}

// out[0..6] = hidden, layers, heads, classes, patch, img, ftype
int vitref_hparams(void *hv, int32_t *out)
{
    vitref_handle *h = (vitref_handle *)hv;
    if (!h) return 1;
    const vit_hparams &hp = h->model.hparams;
    out[0] = hp.hidden_size;
    out[1] = hp.num_hidden_layers;
    out[2] = hp.num_attention_heads;
    out[3] = hp.num_classes;
    out[4] = hp.patch_size;
    out[5] = hp.img_size;
    out[6] = hp.ftype;
    return 0;
}

// One reference forward (vit_predict, vit.cpp:1004-1075) on one pre-processed HWC f32 image.
// probs_out[num_classes] <- state.prediction (post-softmax); logits_out[num_classes] <- the
// captured pre-softmax node.  Either may be NULL.
int vitref_predict(void *hv, const float *img_hwc, int n_threads, float *probs_out, float *logits_out)
{
    vitref_handle *h = (vitref_handle *)hv;
    if (!h || !h->loaded) return 1;
    const vit_hparams &hp = h->model.hparams;
    image_f32 img;
    img.nx = hp.img_size;
    img.ny = hp.img_size;
    img.data.assign(img_hwc, img_hwc + (size_t)3 * hp.img_size * hp.img_size);
    vit_params params;
    params.n_threads = n_threads;
    params.topk = 0; // suppress the label printf loop (vit.cpp:1062)
    std::vector<std::pair<float, int>> predictions;
    int rc = vit_predict_ggml_ref(h->model, h->state, img, params, predictions);
    if (rc != 0) return rc;
    const int nc = hp.num_classes;
    if (probs_out) memcpy(probs_out, ggml_get_data_f32(h->state.prediction), sizeof(float) * nc);
    if (logits_out) memcpy(logits_out, g_logits.data(), sizeof(float) * nc);
    return 0;
}

// Reference preprocessing (vit.cpp:289-305): u8 RGB interleaved -> HWC f32 normalised.
int vitref_preprocess(void *hv, const uint8_t *rgb, int nx, int ny, int bilinear, float *out)
{
    vitref_handle *h = (vitref_handle *)hv;
    if (!h) return 1;
    image_u8 in;
    in.nx = nx;
    in.ny = ny;
    in.data.assign(rgb, rgb + (size_t)3 * nx * ny);
    image_f32 res;
    vit_hparams hp = h->model.hparams;
    hp.interpolation = bilinear ? "bilinear" : "bicubic";
    if (!vit_image_preprocess(in, res, hp)) return 2;
    memcpy(out, res.data.data(), sizeof(float) * res.data.size());
    return 0;
}

// Decode an image file with the reference's stb_image path (vit.cpp:109-127).
// Returns 0 and fills nx/ny; call again with a buffer of 3*nx*ny bytes to get pixels.
int vitref_load_image(const char *path, int *nx, int *ny, uint8_t *rgb_out, int64_t cap)
{
    image_u8 img;
    if (!load_image_from_file(std::string(path), img)) return 1;
    *nx = img.nx;
    *ny = img.ny;
    if (rgb_out)
    {
        if ((int64_t)img.data.size() > cap) return 2;
        memcpy(rgb_out, img.data.data(), img.data.size());
    }
    return 0;
}

void vitref_free(void *hv)
{
    vitref_handle *h = (vitref_handle *)hv;
    if (!h) return;
    if (h->loaded)
    {
        ggml_free(h->state.ctx);
        ggml_free(h->model.ctx);
    }
    delete h;
}

// --- op-level entry points used to pin the restatement's primitives -------------------
// f32 -> f16 -> f32 through the reference's conversion (ggml.c:315-332 / ggml_fp32_to_fp16_row)
void vitref_round_f16(const float *x, float *y, int64_t n)
{
    std::vector<ggml_fp16_t> tmp((size_t)n);
    ggml_fp32_to_fp16_row(x, tmp.data(), (int)n);
    ggml_fp16_to_fp32_row(tmp.data(), y, (int)n);
}


// Run ONE reference ggml op on a [ne1][ne0] f32 matrix (single thread): 0 = ggml_gelu
// (ggml.c:8676-8715), 1 = ggml_soft_max over ne0 (ggml.c:10498-10567), 2 = ggml_norm over ne0
// (ggml.c:8959-9008).  Used to pin the restatement's primitives bit for bit.
int vitref_unary(int op, const float *x, float *y, int64_t ne0, int64_t ne1, float eps)
{
    const size_t nbytes = (size_t)ne0 * ne1 * sizeof(float);
    struct ggml_init_params p = {2 * nbytes + ggml_graph_overhead() + (8u << 20), NULL, false};
    struct ggml_context *ctx = ggml_init(p);
    if (!ctx) return 1;
    struct ggml_tensor *t = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, ne0, ne1);
    memcpy(t->data, x, nbytes);
    struct ggml_tensor *r = op == 0 ? ggml_gelu(ctx, t) : op == 1 ? ggml_soft_max(ctx, t) : ggml_norm(ctx, t, eps);
    struct ggml_cgraph *gf = ggml_new_graph(ctx);
    ggml_build_forward_expand(gf, r);
    ggml_graph_compute_with_ctx(ctx, gf, 1);
    memcpy(y, r->data, nbytes);
    ggml_free(ctx);
    return 0;
}

// y[t][o] = mul_mat(W[o][:], x[t][:]) through the reference's ggml_mul_mat (ggml.c:9388-9597) with
// W stored as wtype (0 f32, 1 f16, 8 q8_0 -- quantised here with the reference quantiser).
int vitref_mul_mat(int wtype, const float *w, const float *x, float *y, int64_t K, int64_t N, int64_t T, int n_threads)
{
    const enum ggml_type ty = wtype == 0 ? GGML_TYPE_F32 : wtype == 1 ? GGML_TYPE_F16 : GGML_TYPE_Q8_0;
    const size_t need = (size_t)(K * N + K * T + N * T) * sizeof(float) * 2 + ggml_graph_overhead() + (16u << 20);
    struct ggml_init_params p = {need, NULL, false};
    struct ggml_context *ctx = ggml_init(p);
    if (!ctx) return 1;
    struct ggml_tensor *tw = ggml_new_tensor_2d(ctx, ty, K, N);
    if (wtype == 0) memcpy(tw->data, w, (size_t)K * N * sizeof(float));
    else if (wtype == 1) ggml_fp32_to_fp16_row(w, (ggml_fp16_t *)tw->data, (int)(K * N));
    else { std::vector<int64_t> hist(16, 0); ggml_quantize_q8_0(w, tw->data, (int)(K * N), (int)K, hist.data()); }
    struct ggml_tensor *tx = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, K, T);
    memcpy(tx->data, x, (size_t)K * T * sizeof(float));
    struct ggml_tensor *r = ggml_mul_mat(ctx, tw, tx);
    struct ggml_cgraph *gf = ggml_new_graph(ctx);
    ggml_build_forward_expand(gf, r);
    ggml_graph_compute_with_ctx(ctx, gf, n_threads);
    memcpy(y, r->data, (size_t)N * T * sizeof(float));
    ggml_free(ctx);
    return 0;
}


// Debug tap: build the reference graph exactly as vit_predict does (vit.cpp:1009-1035), run only its
// first `node_index + 1` nodes, and copy that node's output (valid right after it is computed, before
// ggml-alloc reuses the buffer).  Returns the number of graph nodes; fills ne[4] and op name.

int vitref_tap(void *hv, const float *img_hwc, int n_threads, int node_index, float *out, int64_t cap_floats,
               int64_t *ne, char *op_name, int op_name_cap)
{
    vitref_handle *h = (vitref_handle *)hv;
    if (!h || !h->loaded) return -1;
    const vit_hparams &hp = h->model.hparams;
    image_f32 img;
    img.nx = hp.img_size;
    img.ny = hp.img_size;
    img.data.assign(img_hwc, img_hwc + (size_t)3 * hp.img_size * hp.img_size);
    vit_state &state = h->state;
    static const size_t tensor_alignment = 32;
    state.buf_compute_img_enc.resize(ggml_tensor_overhead() * GGML_DEFAULT_GRAPH_SIZE + ggml_graph_overhead());
    state.allocr = ggml_allocr_new_measure(tensor_alignment);
    struct ggml_cgraph *gf_measure = vit_encode_image_ggml_ref(h->model, state, img);
    size_t alloc_size = ggml_allocr_alloc_graph(state.allocr, gf_measure) + tensor_alignment;
    ggml_allocr_free(state.allocr);
    state.buf_alloc_img_enc.resize(alloc_size);
    state.allocr = ggml_allocr_new(state.buf_alloc_img_enc.data(), state.buf_alloc_img_enc.size(), tensor_alignment);
    ggml_allocr_reset(state.allocr);
    struct ggml_cgraph *gf = vit_encode_image_ggml_ref(h->model, state, img);
    ggml_allocr_alloc_graph(state.allocr, gf);
    const int n_nodes = gf->n_nodes;
    if (node_index >= 0 && node_index < n_nodes)
    {
        gf->n_nodes = node_index + 1;
        ggml_graph_compute_helper(state.work_buffer, gf, n_threads);
        struct ggml_tensor *t = gf->nodes[node_index];
        for (int i = 0; i < 4; ++i) ne[i] = t->ne[i];
        snprintf(op_name, op_name_cap, "%s", ggml_op_name(t->op));
        if (out && t->type == GGML_TYPE_F32 && ggml_is_contiguous(t) && ggml_nelements(t) <= cap_floats)
            memcpy(out, t->data, ggml_nbytes(t));
        else if (out)
            ne[0] = -ne[0]; // signal: not copied (non-contiguous / non-f32 / too large)
    }
    ggml_allocr_free(state.allocr);
    state.allocr = NULL;
    state.work_buffer.clear();
    return n_nodes;
}

} // extern "C"
