// rmatch_kernels.cu — inverse (retain / TopicIndex) match. Placeholder translation unit: the entry points are
// implemented in the next milestone; until then they fail loudly instead of silently doing nothing.
#include <string>

#include "../../include/bfq_gpumatch.h"

extern "C" {
static int32_t nyi() { return BFQ_E_STATE; }
int32_t bfq_rindex_create(int32_t, bfq_rindex**) { return nyi(); }
void bfq_rindex_destroy(bfq_rindex*) {}
int32_t bfq_rindex_reset(bfq_rindex*) { return nyi(); }
int32_t bfq_rindex_add(bfq_rindex*, const uint8_t*, const int64_t*, int32_t, const uint8_t*, const int64_t*, const int32_t*, int64_t, int64_t*) { return nyi(); }
int32_t bfq_rindex_remove(bfq_rindex*, const uint8_t*, int64_t, const uint8_t*, int64_t) { return nyi(); }
int32_t bfq_rindex_commit(bfq_rindex*) { return nyi(); }
int32_t bfq_rindex_lookup(bfq_rindex*, int64_t, uint8_t*, int64_t, int64_t*, uint8_t*, int64_t, int64_t*) { return nyi(); }
int32_t bfq_rmatch(bfq_rindex*, const uint8_t*, const int64_t*, int32_t, const uint8_t*, const int64_t*, const int32_t*, int64_t, const int64_t*, bfq_rresult**) { return nyi(); }
int64_t bfq_rresult_num_filters(const bfq_rresult*) { return 0; }
const int64_t* bfq_rresult_offsets(const bfq_rresult*) { return nullptr; }
const int64_t* bfq_rresult_ids(const bfq_rresult*, int64_t*) { return nullptr; }
const int64_t* bfq_rresult_total_matches(const bfq_rresult*) { return nullptr; }
int32_t bfq_rresult_timings(const bfq_rresult*, double*, int32_t) { return nyi(); }
void bfq_rresult_free(bfq_rresult*) {}
}
