// guiding.cu -- guiding-field kernels and their C-ABI entry points (filled in with the guiding rows).
#include "../../include/b200pg.h"

extern "C" {
int b200pg_train_accumulate(void *) { return -2; }
int b200pg_train_stats_buffer(void *, void **, size_t *) { return -2; }
int b200pg_train_update(void *) { return -2; }
int b200pg_k_grid_lookup(void *, int, const float *, size_t, float *) { return -2; }
int b200pg_k_vmm_pdf_sample(void *, const float *, const float *, const float *, size_t, float *, float *, float *, uint32_t *) { return -2; }
int b200pg_k_bin_samples(void *, const float *, size_t, uint32_t *, uint32_t *, uint32_t *, uint32_t *) { return -2; }
int b200pg_k_em_step(void *, const float *, const float *, const float *, const float *, const float *, size_t) { return -2; }
int b200pg_field_snapshot(void *, float *, size_t *) { return -2; }
int b200pg_field_load(void *, const float *, size_t) { return -2; }
}
