// phi_adapter_testhook.hpp — TEST BUILDS ONLY (-DPHI_ADAPTER_TESTHOOK, oracle/_ref/PHI_gpu_model).
// With PHI_ADAPTER_RESULT_FILE set, run_front_end_result() takes the front end's result from a file instead of calling
// libphi_gpu_index.so, so that the model block of phi_model.hpp can be compared with the unmodified reference's model dump on
// a machine without a GPU (tests/test_model_block.py writes the file from the CPU oracle's anchors, tests/phi_io.py:write_result_file).
// File: "PHIRES3\0", then u64 count_sp_r, n_walks, n_filtered, n_anchors, n_groups, n_group_vtx, member_bytes, then the arrays of
// phi_index_result in declaration order (spectrum, rank_off, group_len, group_vtx, group_member_off, member_walk, minimizers_per_walk,
// anchors_per_walk), each padded to 8 bytes.
#ifndef PHI_ADAPTER_TESTHOOK_HPP
#define PHI_ADAPTER_TESTHOOK_HPP
#include <cstring>
namespace phi_adapter {
inline const phi_index_result *load_result_file(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "Error: cannot open %s\n", path); exit(1); }
    fseek(f, 0, SEEK_END); const long size = ftell(f); fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc(size);                                   // lives as long as the process (test hook)
    if (fread(buf, 1, size, f) != (size_t)size || memcmp(buf, "PHIRES3", 8) != 0) { fprintf(stderr, "Error: bad result file %s\n", path); exit(1); }
    fclose(f);
    const uint64_t *h = (const uint64_t *)(buf + 8);
    phi_index_result *r = (phi_index_result *)calloc(1, sizeof(phi_index_result));
    r->count_sp_r = (int32_t)h[0]; r->n_walks = (uint32_t)h[1]; r->n_filtered = (int64_t)h[2];
    r->n_anchors = h[3]; r->n_groups = h[4]; r->n_group_vtx = h[5];
    const uint64_t mb = h[6];
    char *p = buf + 8 + 7 * 8;
    struct Take { static char *arr(char *&q, uint64_t bytes) { char *a = q; q += (bytes + 7) & ~7ull; return a; } };
    r->spectrum = (const uint64_t *)Take::arr(p, 8 * h[0]);
    r->rank_off = (const uint32_t *)Take::arr(p, 4 * (h[0] + 1));
    r->group_len = (const uint8_t *)Take::arr(p, h[4]);
    r->group_vtx = (const int32_t *)Take::arr(p, 4 * h[5]);
    r->group_member_off = (const uint32_t *)Take::arr(p, 4 * (h[4] + 1));
    if (mb == 2) r->member_walk16 = (const uint16_t *)Take::arr(p, 2 * h[3]); else r->member_walk32 = (const int32_t *)Take::arr(p, 4 * h[3]);
    r->minimizers_per_walk = (const uint64_t *)Take::arr(p, 8 * h[1]);
    r->anchors_per_walk = (const uint64_t *)Take::arr(p, 8 * h[1]);
    if (p - buf != size) { fprintf(stderr, "Error: result file %s has %ld bytes, expected %ld\n", path, size, (long)(p - buf)); exit(1); }
    return r;
}
}  // namespace phi_adapter
#endif
