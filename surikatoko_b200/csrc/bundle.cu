// suriko-b200 -- bundle files: binary dump / load of the flat BA problem (include/srk/bundle_c_api.h).  Host code only.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/srk/bundle_c_api.h"

extern "C" void srk_internal_set_error(const char* s);   // engine.cu: the string behind srk_last_error()

namespace {

const char kMagic[8] = {'S', 'R', 'K', 'B', 'N', 'D', 'L', '1'};

struct Header { char magic[8]; int64_t n_cams, n_points, n_obs; int32_t shared_K, reserved; double f0; };
static_assert(sizeof(Header) == 48, "bundle header is 48 bytes, no padding");

struct Fnv {
    uint64_t h = 1469598103934665603ull;
    void add(const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } }
};

int fail(const std::string& s, int code = SRK_E_INVALID_ARG) { srk_internal_set_error(s.c_str()); return code; }

struct File {
    FILE* f = nullptr;
    ~File() { if (f != nullptr) fclose(f); }
};

bool sizes_ok(const Header& h) {
    return h.n_cams >= 0 && h.n_points >= 0 && h.n_obs >= 0 && h.n_cams < (1ll << 31) && h.n_points < (1ll << 31) && h.n_obs < (1ll << 40) &&
           (h.shared_K == 0 || h.shared_K == 1);
}
int64_t payload_bytes(const Header& h) {
    return 4 * h.n_obs * 2 + 16 * h.n_obs + 24 * h.n_points + 96 * h.n_cams + 72 * (h.shared_K ? 1 : h.n_cams);
}
int read_header(File& fl, const char* path, Header& h) {
    fl.f = fopen(path, "rb");
    if (fl.f == nullptr) return fail(std::string("Can't open file ") + path);
    if (fread(&h, sizeof(h), 1, fl.f) != 1) return fail(std::string("bundle file is shorter than its header: ") + path);
    if (memcmp(h.magic, kMagic, 8) != 0) return fail(std::string("not a bundle file (bad magic): ") + path);
    if (!sizes_ok(h)) return fail(std::string("bundle header holds impossible sizes: ") + path);
    return SRK_OK;
}

}  // namespace

extern "C" {

int srk_bundle_write(const char* path, const srk_ba_problem* p) {
    if (path == nullptr || p == nullptr) return fail("null argument");
    Header h; memcpy(h.magic, kMagic, 8);
    h.n_cams = p->n_cams; h.n_points = p->n_points; h.n_obs = p->n_obs; h.shared_K = p->shared_K ? 1 : 0; h.reserved = 0; h.f0 = p->f0;
    if (!sizes_ok(h)) return fail("problem sizes out of range");
    if ((p->n_obs > 0 && (p->obs_cam == nullptr || p->obs_point == nullptr || p->obs_xy == nullptr)) || (p->n_points > 0 && p->points == nullptr) ||
        (p->n_cams > 0 && p->cams == nullptr) || p->K == nullptr)
        return fail("null array in the problem");
    File fl; fl.f = fopen(path, "wb");
    if (fl.f == nullptr) return fail(std::string("Can't open file for writing ") + path);
    Fnv sum;
    auto put = [&](const void* d, size_t bytes) -> bool { if (bytes == 0) return true; sum.add(d, bytes); return fwrite(d, 1, bytes, fl.f) == bytes; };
    bool ok = put(&h, sizeof(h)) && put(p->obs_cam, 4 * (size_t)h.n_obs) && put(p->obs_point, 4 * (size_t)h.n_obs) && put(p->obs_xy, 16 * (size_t)h.n_obs) &&
              put(p->points, 24 * (size_t)h.n_points) && put(p->cams, 96 * (size_t)h.n_cams) && put(p->K, 72 * (size_t)(h.shared_K ? 1 : h.n_cams));
    const uint64_t s = sum.h;
    ok = ok && fwrite(&s, 8, 1, fl.f) == 1;
    if (!ok) return fail(std::string("write failed: ") + path);
    return SRK_OK;
}

int srk_bundle_read_header(const char* path, int64_t* n_cams, int64_t* n_points, int64_t* n_obs, int32_t* shared_K, double* f0) {
    if (path == nullptr) return fail("null argument");
    File fl; Header h;
    const int rc = read_header(fl, path, h);
    if (rc != SRK_OK) return rc;
    if (n_cams != nullptr) *n_cams = h.n_cams;
    if (n_points != nullptr) *n_points = h.n_points;
    if (n_obs != nullptr) *n_obs = h.n_obs;
    if (shared_K != nullptr) *shared_K = h.shared_K;
    if (f0 != nullptr) *f0 = h.f0;
    return SRK_OK;
}

int srk_bundle_read(const char* path, int32_t* obs_cam, int32_t* obs_point, double* obs_xy, double* points, double* cams, double* K) {
    if (path == nullptr) return fail("null argument");
    File fl; Header h;
    const int rc = read_header(fl, path, h);
    if (rc != SRK_OK) return rc;
    if ((h.n_obs > 0 && (obs_cam == nullptr || obs_point == nullptr || obs_xy == nullptr)) || (h.n_points > 0 && points == nullptr) || (h.n_cams > 0 && cams == nullptr) ||
        K == nullptr)
        return fail("null output array");
    Fnv sum; sum.add(&h, sizeof(h));
    auto get = [&](void* d, size_t bytes) -> bool { if (bytes == 0) return true; if (fread(d, 1, bytes, fl.f) != bytes) return false; sum.add(d, bytes); return true; };
    const bool ok = get(obs_cam, 4 * (size_t)h.n_obs) && get(obs_point, 4 * (size_t)h.n_obs) && get(obs_xy, 16 * (size_t)h.n_obs) && get(points, 24 * (size_t)h.n_points) &&
                    get(cams, 96 * (size_t)h.n_cams) && get(K, 72 * (size_t)(h.shared_K ? 1 : h.n_cams));
    uint64_t stored = 0;
    if (!ok || fread(&stored, 8, 1, fl.f) != 1)
        return fail("bundle file is truncated: " + std::string(path) + " (payload " + std::to_string(payload_bytes(h)) + " bytes expected)");
    if (stored != sum.h) return fail(std::string("bundle checksum mismatch: ") + path);
    char extra;
    if (fread(&extra, 1, 1, fl.f) == 1) return fail(std::string("bundle file has trailing bytes: ") + path);
    return SRK_OK;
}

}  // extern "C"
