// dist.cu -- sharded state across the GPUs of one box (placeholder until the single-GPU path is green)
#include "../../include/cunqa_b200.h"
#include <string>
extern "C" {
int cb200_dist_unique_id(uint8_t *) { return CB200_ERR_UNSUPPORTED; }
int cb200_state_create_dist(cb200_state **, int, int, int, int, int, const uint8_t *) { return CB200_ERR_UNSUPPORTED; }
int cb200_dist_export_handle(cb200_state *, uint8_t *) { return CB200_ERR_UNSUPPORTED; }
int cb200_dist_import_handles(cb200_state *, const uint8_t *) { return CB200_ERR_UNSUPPORTED; }
}
