/* Plain-C client of the rpo_b200 C ABI (include/rpo_b200.h): shows that the boundary needs neither Python nor
 * torch -- any host language with a C FFI binds it the same way.  Without a GPU it can only exercise argument
 * validation; tests/test_cabi.py compiles, links and runs it.
 *   gcc -std=c99 -Iinclude integration/c_client.c -Lrpo_b200/lib -lrpo_b200 -Wl,-rpath,$PWD/rpo_b200/lib */
#include <stdio.h>
#include <string.h>

#include "rpo_b200.h"

int main(void) {
  RpoConfig cfg;
  RpoHandle *h = NULL;
  memset(&cfg, 0, sizeof cfg);
  cfg.dtype = 7; /* not an RPO_* element type */
  if (rpo_create(&cfg, &h) != RPO_ERR_INVALID || h != NULL) return 1;
  if (strstr(rpo_last_error(), "dtype") == NULL) return 2;
  cfg.dtype = RPO_F16;
  cfg.K = 0; /* trainers/rpo.py:47 asserts K >= 1 */
  cfg.n_cls = 2; cfg.ctx_len = 77; cfg.embed_dim = 512; cfg.max_batch = 1;
  cfg.v_width = 768; cfg.v_layers = 12; cfg.v_heads = 12; cfg.v_patch = 16; cfg.v_res = 224;
  cfg.t_width = 512; cfg.t_layers = 12; cfg.t_heads = 8;
  if (rpo_create(&cfg, &h) != RPO_ERR_INVALID) return 3;
  if (rpo_backward(NULL, NULL, NULL) != RPO_ERR_INVALID) return 4;
  if (rpo_forward_text(NULL, NULL, NULL) != RPO_ERR_INVALID) return 5;
  printf("rpo_b200 C ABI version %d: argument validation ok (%s)\n", rpo_version(), rpo_last_error());
  return 0;
}
