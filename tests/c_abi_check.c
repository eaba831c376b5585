/* Plain-C client of libnpe_cuda.so: the header compiles as C, the library links from C, errors come back as status codes
 * and messages (never an abort), and -- when a GPU is present -- a tiny batch runs forward / backward / optimiser.
 * Built and run by tests/test_abi.py.  Exit code 0 = all checks passed; prints "gpu" or "nogpu". */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/npe_cuda.h"

#define CHECK(c, msg) do { if (!(c)) { fprintf(stderr, "FAIL: %s (%s:%d)\n", msg, __FILE__, __LINE__); return 1; } } while (0)

int main(void) {
    npe_config cfg;
    npe_handle* h = NULL;
    CHECK(npe_abi_version() == NPE_ABI_VERSION, "abi version");
    npe_default_config(&cfg);
    CHECK(cfg.struct_size == (int32_t)sizeof(npe_config) && cfg.rnn_dim == 50 && cfg.layers == 5 && cfg.max_ctx == 12, "defaults");
    CHECK(npe_create(NULL, &h) == NPE_ERR_INVALID, "null config is rejected");
    cfg.rnn_dim = 24;                                   /* the -debug shape is outside the north-star path */
    int rc = npe_create(&cfg, &h);
    CHECK(rc == NPE_ERR_UNSUPPORTED || rc == NPE_ERR_CUDA, "unsupported width is rejected");
    CHECK(strlen(npe_last_error(NULL)) > 0, "error message");
    cfg.rnn_dim = 50;
    rc = npe_create(&cfg, &h);
    if (rc != NPE_OK) {                                 /* no device: fails loudly, there is no CPU path */
        CHECK(rc == NPE_ERR_CUDA && h == NULL, "create without a GPU fails with NPE_ERR_CUDA");
        CHECK(strstr(npe_last_error(NULL), "CUDA") != NULL, "message names CUDA");
        printf("nogpu\n");
        return 0;
    }
    CHECK(npe_param_count(h) == NPE_NUM_PARAMS, "parameter count");
    float* p = (float*)malloc(sizeof(float) * NPE_NUM_PARAMS);
    for (int i = 0; i < NPE_NUM_PARAMS; ++i) p[i] = 0.01f * (float)((i * 37) % 19 - 9);
    CHECK(npe_params_set(h, p, NPE_NUM_PARAMS) == NPE_OK, "params_set");
    CHECK(npe_backward(h, 0) == NPE_ERR_STATE, "backward before forward is a state error");
    /* two foci, one context each, 0.1 apart (inside the 210 px neighbourhood) */
    float this_past[2 * 44] = {0}, ctx[2 * 44] = {0}, y[2 * 22] = {0}, pred[2 * 22], loss3[3];
    for (int b = 0; b < 2; ++b)
        for (int f = 0; f < 2; ++f) {
            float* a = this_past + b * 44 + f * 22; float* c = ctx + b * 44 + f * 22;
            a[0] = 0.3f + 0.01f * f; a[1] = 0.3f; a[2] = 0.1f; a[6] = 1.f; a[10] = 1.f; a[14] = 1.f; a[16] = a[18] = a[20] = 1.f;
            memcpy(c, a, 22 * sizeof(float)); c[0] += 0.1f; c[2] = -0.1f;
        }
    y[2] = 0.05f; y[22 + 3] = -0.05f;
    CHECK(npe_batch_upload(h, this_past, ctx, y, 2, 1) == NPE_OK, "batch_upload");
    CHECK(npe_forward(h, pred, loss3) == NPE_OK, "forward");
    CHECK(isfinite(loss3[0]) && loss3[0] > 0.f && isfinite(pred[0]) && pred[6] > 0.f && pred[6] < 1.f, "finite outputs, sigmoid columns");
    CHECK(npe_backward(h, 0) == NPE_OK, "backward");
    float* g = (float*)malloc(sizeof(float) * NPE_NUM_PARAMS);
    CHECK(npe_grads_get(h, g, NPE_NUM_PARAMS) == NPE_OK, "grads_get");
    double n2 = 0; for (int i = 0; i < NPE_NUM_PARAMS; ++i) n2 += (double)g[i] * g[i];
    CHECK(n2 > 0 && isfinite(n2), "non-zero finite gradient");
    CHECK(npe_adam_step(h, 3e-4f, 0.9f, 0.999f, 1e-8f) == NPE_OK && npe_params_get(h, g, NPE_NUM_PARAMS) == NPE_OK, "adam");
    int moved = 0; for (int i = 0; i < NPE_NUM_PARAMS; ++i) moved += fabsf(g[i] - p[i]) > 1e-6f;
    CHECK(moved > 1000, "parameters moved");
    CHECK(npe_rollout(h, 0, 4, 1, 0, NULL, NULL) == NPE_ERR_STATE, "rollout needs scenes");
    CHECK(npe_destroy(h) == NPE_OK, "destroy");
    free(p); free(g);
    printf("gpu\n");
    return 0;
}
