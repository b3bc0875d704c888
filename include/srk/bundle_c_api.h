/* suriko-b200 -- bundle files: a binary dump of the flat problem srk_ba_solve takes, so that a large
 * synthetic scene is generated once and shared between the CPU oracle, the GPU engine and a reviewer.  The reference has no such
 * format (its scenes live in FragmentMap / CornerTrackRepository objects, obs-geom.h:199-304); the file holds exactly what those
 * containers flatten to through GetCorner / EachCorner (include/srk/ba_c_api.h, srk_ba_problem).
 *
 * Layout, little-endian, no padding:
 *   char     magic[8]  = "SRKBNDL1"
 *   int64    n_cams, n_points, n_obs
 *   int32    shared_K, reserved (0)
 *   double   f0
 *   int32    obs_cam[n_obs]       int32  obs_point[n_obs]      double obs_xy[2 n_obs]
 *   double   points[3 n_points]   double cams[12 n_cams]       double K[9] or K[9 n_cams]
 *   uint64   FNV-1a 64 of every byte before it
 * Host code only; plain pointers and sizes; 0 = ok, negative = SRK_E_* with the text in srk_last_error(). */
#ifndef SRK_BUNDLE_C_API_H
#define SRK_BUNDLE_C_API_H
#include <stdint.h>
#include "ba_c_api.h"
#ifdef __cplusplus
extern "C" {
#endif

SRK_API int srk_bundle_write(const char* path, const srk_ba_problem* problem);
/* sizes of the scene in `path` (any output pointer may be null) */
SRK_API int srk_bundle_read_header(const char* path, int64_t* n_cams, int64_t* n_points, int64_t* n_obs, int32_t* shared_K, double* f0);
/* fills caller-allocated arrays of the sizes the header announces; verifies magic, sizes, file length and the checksum */
SRK_API int srk_bundle_read(const char* path, int32_t* obs_cam, int32_t* obs_point, double* obs_xy, double* points, double* cams, double* K);

#ifdef __cplusplus
}
#endif
#endif /* SRK_BUNDLE_C_API_H */
