/*
 * pcdet_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (pcdet_b200/) never links, imports or calls it.
 *
 * What is restated, and from where:
 *   - voxelisation, rulebook and gather-GEMM-scatter live in the third-party package `spconv`
 *     v1.0 @ 8da6f967fb9a054d8870c3515b1b44eca2103634 (pinned by /root/reference README.md:63 and
 *     docker/Dockerfile:104-108).  That package is NOT vendored in /root/reference and cannot be
 *     installed here, so those functions restate its published algorithm as recorded in
 *     SURVEY.md Appendix A.1/A.3/A.4 and are anchored on the reference call sites
 *     (pcdet/datasets/dataset.py:163-181, pcdet/models/rpn/rpn_backbone.py:7-103).
 *     PARITY UNPINNED for these three: the reference holds no golden vectors or tests for them.
 *   - rotated BEV overlap / IoU / NMS restate pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:35-293
 *     and the host sweep of pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:79-126.  These are pinned against
 *     the reference's own kernel compiled into oracle/_ref/ (see oracle/build_ref.sh) through the
 *     fixtures in tests/golden/.
 *
 * Plain C99, single thread, no dependencies.  Build: oracle/build.sh (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * A.1  points_to_voxel  (spconv v1.0 points_to_voxel_3d_np<float,3>, reverse index -> zyx)
 *
 * coor_to_voxelidx: caller-provided dense int32 lookup of grid[2]*grid[1]*grid[0] cells, all -1 on
 * entry, restored to -1 on exit (the reference keeps one per VoxelGenerator object).
 * overflow_break != 0 -> stop at the first point that would open voxel number max_voxels (v1.0);
 * 0 -> skip that point and keep filling existing voxels (v1.1+).
 * pt_idx (optional, may be NULL): (max_voxels, max_points) original point index per slot, -1 padded
 * (the fork's `voxel_pt_indices_into_original_pt_cloud`, pcdet/experiments.py:236-241).
 * Returns the number of voxels.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_points_to_voxel(const float *points, int n_points, int n_feat,
                                const float *voxel_size, const float *range, const int *grid_xyz,
                                int max_points, int max_voxels, int overflow_break,
                                int32_t *coor_to_voxelidx, float *voxels, int32_t *coors_zyx,
                                int32_t *num_points_per_voxel, int32_t *pt_idx)
{
    int voxel_num = 0;
    const int gx = grid_xyz[0], gy = grid_xyz[1], gz = grid_xyz[2];
    (void)gz;
    for (int i = 0; i < n_points; ++i) {
        int c[3];
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            /* fp32 subtract, fp32 IEEE divide, floor -- never a reciprocal multiply */
            float q = (points[(size_t)i * n_feat + j] - range[j]) / voxel_size[j];
            int cj = (int)floorf(q);
            if (cj < 0 || cj >= grid_xyz[j]) { failed = 1; break; }
            c[j] = cj;
        }
        if (failed) continue;
        const size_t cell = ((size_t)c[2] * gy + c[1]) * gx + c[0];
        int vid = coor_to_voxelidx[cell];
        if (vid == -1) {
            vid = voxel_num;
            if (voxel_num >= max_voxels) {
                if (overflow_break) break;
                continue;
            }
            voxel_num += 1;
            coor_to_voxelidx[cell] = vid;
            coors_zyx[vid * 3 + 0] = c[2];
            coors_zyx[vid * 3 + 1] = c[1];
            coors_zyx[vid * 3 + 2] = c[0];
        }
        int n = num_points_per_voxel[vid];
        if (n < max_points) {
            memcpy(voxels + ((size_t)vid * max_points + n) * n_feat, points + (size_t)i * n_feat,
                   sizeof(float) * n_feat);
            if (pt_idx) pt_idx[(size_t)vid * max_points + n] = i;
            num_points_per_voxel[vid] = n + 1;
        }
    }
    for (int v = 0; v < voxel_num; ++v) {
        const size_t cell = ((size_t)coors_zyx[v * 3] * gy + coors_zyx[v * 3 + 1]) * gx + coors_zyx[v * 3 + 2];
        coor_to_voxelidx[cell] = -1;
    }
    return voxel_num;
}

/* pcdet/models/vfe/vfe_utils.py:26-34: sum over the point axis / num_points (zero padded slots). */
ORC_API void orc_vfe_mean(const float *voxels, const int32_t *num_points, int n_vox, int max_points,
                          int n_feat, float *out)
{
    for (int v = 0; v < n_vox; ++v)
        for (int c = 0; c < n_feat; ++c) {
            float s = 0.f;
            for (int p = 0; p < max_points; ++p) s += voxels[((size_t)v * max_points + p) * n_feat + c];
            out[(size_t)v * n_feat + c] = s / (float)num_points[v];
        }
}

/* ------------------------------------------------------------------------------------------
 * A.3  get_indice_pairs  (spconv v1.0 getIndicePair<3>, CPU path: first-touch output order)
 * ---------------------------------------------------------------------------------------- */

/* getValidOutPos: every (output position, kernel offset) an input voxel contributes to.
 * C integer division (truncation) exactly as upstream.  Returns the count; out_pos is (n,3),
 * out_off is the row-major offset index over (kz,ky,kx). */
static int valid_out_pos(const int *in_zyx, const int *ksize, const int *stride, const int *pad,
                         const int *dil, const int *out_shape, int *out_pos, int *out_off)
{
    int lo[3], hi[3], cnt[3];
    int total = 1;
    for (int d = 0; d < 3; ++d) {
        lo[d] = (in_zyx[d] - (ksize[d] - 1) * dil[d] - 1 + stride[d] + pad[d]) / stride[d];
        hi[d] = (in_zyx[d] + pad[d]) / stride[d];
        cnt[d] = (hi[d] - lo[d]) / dil[d] + 1;
        total *= cnt[d];
    }
    int n = 0;
    for (int idx = 0; idx < total; ++idx) {
        int valid = 1, m = 1, off = 0, rem = idx, pos[3];
        for (int d = 2; d >= 0; --d) {
            int c = rem % cnt[d];
            rem /= cnt[d];
            int val = hi[d] - c * dil[d];
            if (val < 0 || val > out_shape[d] - 1) valid = 0;
            pos[d] = val;
            off += m * ((in_zyx[d] - val * stride[d] + pad[d]) / dil[d]);
            m *= ksize[d];
        }
        if (valid) {
            out_pos[n * 3 + 0] = pos[0]; out_pos[n * 3 + 1] = pos[1]; out_pos[n * 3 + 2] = pos[2];
            out_off[n] = off;
            ++n;
        }
    }
    return n;
}

/* Strided (regular) sparse convolution rulebook.
 * indices (n_in,4) [b,z,y,x]; grid: caller-provided int32 of batch*prod(out_shape) cells, holding 0
 * on entry ("value+1" encoding so calloc pages stay untouched) and restored on exit.
 * pairs (K,2,n_in) int32 pre-filled with -1; pair_num (K) zeroed; out_ids (cap,4).
 * Returns n_out. */
ORC_API int orc_rulebook_conv(const int32_t *indices, int n_in, int batch, const int *spatial_shape,
                              const int *out_shape, const int *ksize, const int *stride,
                              const int *pad, const int *dil, int32_t *grid, int32_t *pairs,
                              int32_t *pair_num, int32_t *out_ids)
{
    (void)batch; (void)spatial_shape;
    const int K = ksize[0] * ksize[1] * ksize[2];
    const size_t vol = (size_t)out_shape[0] * out_shape[1] * out_shape[2];
    int n_out = 0;
    int out_pos[27 * 3 * 8], out_off[27 * 8];
    (void)K;
    for (int j = 0; j < n_in; ++j) {
        const int32_t *row = indices + (size_t)j * 4;
        int in_zyx[3] = {row[1], row[2], row[3]};
        int n = valid_out_pos(in_zyx, ksize, stride, pad, dil, out_shape, out_pos, out_off);
        for (int t = 0; t < n; ++t) {
            const int *p = out_pos + t * 3;
            size_t cell = (size_t)row[0] * vol + ((size_t)p[0] * out_shape[1] + p[1]) * out_shape[2] + p[2];
            int oid;
            if (grid[cell] == 0) {
                oid = n_out++;
                grid[cell] = oid + 1;
                out_ids[oid * 4 + 0] = row[0];
                out_ids[oid * 4 + 1] = p[0]; out_ids[oid * 4 + 2] = p[1]; out_ids[oid * 4 + 3] = p[2];
            } else {
                oid = grid[cell] - 1;
            }
            int off = out_off[t];
            int slot = pair_num[off]++;
            pairs[((size_t)off * 2 + 0) * n_in + slot] = j;
            pairs[((size_t)off * 2 + 1) * n_in + slot] = oid;
        }
    }
    for (int o = 0; o < n_out; ++o) {
        const int32_t *r = out_ids + (size_t)o * 4;
        size_t cell = (size_t)r[0] * vol + ((size_t)r[1] * out_shape[1] + r[2]) * out_shape[2] + r[3];
        grid[cell] = 0;
    }
    return n_out;
}

/* Submanifold rulebook: stride forced to 1 and padding to k/2 (A.3), outputs == inputs. */
ORC_API int orc_rulebook_subm(const int32_t *indices, int n_in, int batch, const int *spatial_shape,
                              const int *ksize, const int *dil, int32_t *grid, int32_t *pairs,
                              int32_t *pair_num)
{
    (void)batch;
    const size_t vol = (size_t)spatial_shape[0] * spatial_shape[1] * spatial_shape[2];
    int stride[3] = {1, 1, 1}, pad[3];
    for (int d = 0; d < 3; ++d) pad[d] = ksize[d] / 2;
    int out_pos[27 * 3 * 8], out_off[27 * 8];
    for (int j = 0; j < n_in; ++j) {
        const int32_t *r = indices + (size_t)j * 4;
        size_t cell = (size_t)r[0] * vol + ((size_t)r[1] * spatial_shape[1] + r[2]) * spatial_shape[2] + r[3];
        grid[cell] = j + 1;
    }
    for (int j = 0; j < n_in; ++j) {
        const int32_t *row = indices + (size_t)j * 4;
        int in_zyx[3] = {row[1], row[2], row[3]};
        int n = valid_out_pos(in_zyx, ksize, stride, pad, dil, spatial_shape, out_pos, out_off);
        for (int t = 0; t < n; ++t) {
            const int *p = out_pos + t * 3;
            size_t cell = (size_t)row[0] * vol + ((size_t)p[0] * spatial_shape[1] + p[1]) * spatial_shape[2] + p[2];
            if (grid[cell] > 0) {
                int off = out_off[t];
                int slot = pair_num[off]++;
                pairs[((size_t)off * 2 + 0) * n_in + slot] = j;
                pairs[((size_t)off * 2 + 1) * n_in + slot] = grid[cell] - 1;
            }
        }
    }
    for (int j = 0; j < n_in; ++j) {
        const int32_t *r = indices + (size_t)j * 4;
        size_t cell = (size_t)r[0] * vol + ((size_t)r[1] * spatial_shape[1] + r[2]) * spatial_shape[2] + r[3];
        grid[cell] = 0;
    }
    return n_in;
}

/* ------------------------------------------------------------------------------------------
 * A.4  indice_conv forward.  out (n_out, c_out) zero on entry.
 * subm: the offset with the most pairs (the centre) is applied as a plain features @ W product,
 * then the remaining offsets in ascending order; inverse swaps the pair roles.
 * acc64 != 0 accumulates in double (the "truth" used for tolerance tests).
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_indice_conv(const float *features, const float *filters, const int32_t *pairs,
                             const int32_t *pair_num, int n_in_cap, int n_out, int c_in, int c_out,
                             int K, int subm, int inverse, int acc64, float *out)
{
    int centre = -1;
    double *out64 = NULL;
    if (acc64) out64 = (double *)calloc((size_t)n_out * c_out, sizeof(double));
    if (subm) {
        int best = -1;
        for (int k = 0; k < K; ++k) if (pair_num[k] > best) { best = pair_num[k]; centre = k; }
        const float *w = filters + (size_t)centre * c_in * c_out;
        for (int r = 0; r < n_out; ++r)
            for (int co = 0; co < c_out; ++co) {
                if (acc64) {
                    double s = 0;
                    for (int ci = 0; ci < c_in; ++ci) s += (double)features[(size_t)r * c_in + ci] * w[ci * c_out + co];
                    out64[(size_t)r * c_out + co] = s;
                } else {
                    float s = 0;
                    for (int ci = 0; ci < c_in; ++ci) s += features[(size_t)r * c_in + ci] * w[ci * c_out + co];
                    out[(size_t)r * c_out + co] = s;
                }
            }
    }
    for (int k = 0; k < K; ++k) {
        if (k == centre || pair_num[k] <= 0) continue;
        const float *w = filters + (size_t)k * c_in * c_out;
        const int32_t *pin = pairs + ((size_t)k * 2 + (inverse ? 1 : 0)) * n_in_cap;
        const int32_t *pout = pairs + ((size_t)k * 2 + (inverse ? 0 : 1)) * n_in_cap;
        for (int t = 0; t < pair_num[k]; ++t) {
            const float *f = features + (size_t)pin[t] * c_in;
            for (int co = 0; co < c_out; ++co) {
                if (acc64) {
                    double s = 0;
                    for (int ci = 0; ci < c_in; ++ci) s += (double)f[ci] * w[ci * c_out + co];
                    out64[(size_t)pout[t] * c_out + co] += s;
                } else {
                    float s = 0;
                    for (int ci = 0; ci < c_in; ++ci) s += f[ci] * w[ci * c_out + co];
                    out[(size_t)pout[t] * c_out + co] += s;
                }
            }
        }
    }
    if (acc64) {
        for (size_t i = 0; i < (size_t)n_out * c_out; ++i) out[i] = (float)out64[i];
        free(out64);
    }
}

/* ------------------------------------------------------------------------------------------
 * Rotated BEV overlap / IoU / NMS  (iou3d_nms_kernel.cu:35-293, iou3d_nms.cpp:79-126)
 * Boxes are (x1, y1, x2, y2, ry); all arithmetic fp32 like the device code.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float x, y; } pt_t;
static const float ORC_EPS = 1e-8f;

static float cross3(pt_t p1, pt_t p2, pt_t p0)
{
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

/* iou3d_nms_kernel.cu:43-49 */
static int bbox_touch(pt_t p1, pt_t p2, pt_t q1, pt_t q2)
{
    return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
           fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}

/* iou3d_nms_kernel.cu:51-67 */
static int corner_inside(const float *box, pt_t p)
{
    const float margin = 1e-5f;
    float cx = (box[0] + box[2]) / 2, cy = (box[1] + box[3]) / 2;
    float ca = cosf(-box[4]), sa = sinf(-box[4]);
    float rx = (p.x - cx) * ca + (p.y - cy) * sa + cx;
    float ry = -(p.x - cx) * sa + (p.y - cy) * ca + cy;
    return rx > box[0] - margin && rx < box[2] + margin && ry > box[1] - margin && ry < box[3] + margin;
}

/* iou3d_nms_kernel.cu:69-98 : proper crossing of segment p0p1 with q0q1 */
static int seg_cross(pt_t p1, pt_t p0, pt_t q1, pt_t q0, pt_t *ans)
{
    if (!bbox_touch(p0, p1, q0, q1)) return 0;
    float s1 = cross3(q0, p1, p0);
    float s2 = cross3(p1, q1, p0);
    float s3 = cross3(p0, q1, q0);
    float s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > ORC_EPS) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static void spin(pt_t c, float ca, float sa, pt_t *p)
{
    float nx = (p->x - c.x) * ca + (p->y - c.y) * sa + c.x;
    float ny = -(p->x - c.x) * sa + (p->y - c.y) * ca + c.y;
    p->x = nx; p->y = ny;
}

/* iou3d_nms_kernel.cu:109-212 */
ORC_API float orc_box_overlap(const float *a, const float *b)
{
    pt_t ca_ = {(a[0] + a[2]) / 2, (a[1] + a[3]) / 2};
    pt_t cb_ = {(b[0] + b[2]) / 2, (b[1] + b[3]) / 2};
    pt_t A[5] = {{a[0], a[1]}, {a[2], a[1]}, {a[2], a[3]}, {a[0], a[3]}};
    pt_t B[5] = {{b[0], b[1]}, {b[2], b[1]}, {b[2], b[3]}, {b[0], b[3]}};
    float cosa = cosf(a[4]), sina = sinf(a[4]), cosb = cosf(b[4]), sinb = sinf(b[4]);
    for (int k = 0; k < 4; ++k) { spin(ca_, cosa, sina, &A[k]); spin(cb_, cosb, sinb, &B[k]); }
    A[4] = A[0]; B[4] = B[0];

    pt_t poly[16], centre = {0, 0};
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_cross(A[i + 1], A[i], B[j + 1], B[j], &poly[cnt])) {
                centre.x += poly[cnt].x; centre.y += poly[cnt].y; ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (corner_inside(a, B[k])) { centre.x += B[k].x; centre.y += B[k].y; poly[cnt++] = B[k]; }
        if (corner_inside(b, A[k])) { centre.x += A[k].x; centre.y += A[k].y; poly[cnt++] = A[k]; }
    }
    centre.x /= cnt; centre.y /= cnt;
    /* bubble sort by polar angle about the centroid (iou3d_nms_kernel.cu:189-199) */
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(poly[i].y - centre.y, poly[i].x - centre.x) >
                atan2f(poly[i + 1].y - centre.y, poly[i + 1].x - centre.x)) {
                pt_t t = poly[i]; poly[i] = poly[i + 1]; poly[i + 1] = t;
            }
    float area = 0;
    for (int k = 0; k < cnt - 1; ++k) {
        pt_t u = {poly[k].x - poly[0].x, poly[k].y - poly[0].y};
        pt_t v = {poly[k + 1].x - poly[0].x, poly[k + 1].y - poly[0].y};
        area += u.x * v.y - u.y * v.x;
    }
    return fabsf(area) / 2.0f;
}

/* iou3d_nms_kernel.cu:214-222 */
ORC_API float orc_iou_bev(const float *a, const float *b)
{
    float sa = (a[2] - a[0]) * (a[3] - a[1]);
    float sb = (b[2] - b[0]) * (b[3] - b[1]);
    float so = orc_box_overlap(a, b);
    return so / fmaxf(sa + sb - so, ORC_EPS);
}

/* iou3d_nms_kernel.cu:296-305 */
ORC_API float orc_iou_normal(const float *a, const float *b)
{
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    float inter = w * h;
    float sa = (a[2] - a[0]) * (a[3] - a[1]);
    float sb = (b[2] - b[0]) * (b[3] - b[1]);
    return inter / fmaxf(sa + sb - inter, ORC_EPS);
}

ORC_API void orc_boxes_overlap_bev(const float *a, int na, const float *b, int nb, float *out)
{
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = orc_box_overlap(a + i * 5, b + j * 5);
}

ORC_API void orc_boxes_iou_bev(const float *a, int na, const float *b, int nb, float *out)
{
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = orc_iou_bev(a + i * 5, b + j * 5);
}

/* Greedy suppression over score-sorted boxes.  Equivalent to the mask + host sweep of
 * iou3d_nms.cpp:79-126: box i is kept iff no earlier KEPT box j<i has iou(j,i) > thresh.
 * normal != 0 uses the axis-aligned IoU (nms_normal_gpu).  Returns the keep count. */
ORC_API int orc_nms(const float *boxes, int n, float thresh, int normal, int64_t *keep)
{
    unsigned char *dead = (unsigned char *)calloc((size_t)n + 1, 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[nk++] = i;
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            float v = normal ? orc_iou_normal(boxes + i * 5, boxes + j * 5)
                             : orc_iou_bev(boxes + i * 5, boxes + j * 5);
            if (v > thresh) dead[j] = 1;
        }
    }
    free(dead);
    return nk;
}

/* ------------------------------------------------------------------------------------------
 * Exact-geometry helper (NOT part of the reference): fp64 convex clipping of rectangle A by
 * rectangle B.  Used only by the test generators to find pairs whose IoU sits within a margin of
 * the threshold, where any two fp32 implementations may legitimately disagree.
 * ---------------------------------------------------------------------------------------- */
static void rect_corners64(const float *b, double *xs, double *ys)
{
    double cx = ((double)b[0] + b[2]) / 2, cy = ((double)b[1] + b[3]) / 2;
    double hx = ((double)b[2] - b[0]) / 2, hy = ((double)b[3] - b[1]) / 2;
    double c = cos((double)b[4]), s = sin((double)b[4]);
    const double lx[4] = {-hx, hx, hx, -hx}, ly[4] = {-hy, -hy, hy, hy};
    for (int k = 0; k < 4; ++k) {
        /* same sense of rotation as rotate_around_center (iou3d_nms_kernel.cu:100-104) */
        xs[k] = lx[k] * c + ly[k] * s + cx;
        ys[k] = -lx[k] * s + ly[k] * c + cy;
    }
}

ORC_API double orc_box_overlap64(const float *a, const float *b)
{
    double px[16], py[16], qx[16], qy[16], bx[4], by[4];
    int n = 4;
    rect_corners64(a, px, py);
    rect_corners64(b, bx, by);
    /* orientation of B */
    double orient = 0;
    for (int k = 0; k < 4; ++k) orient += bx[k] * by[(k + 1) % 4] - bx[(k + 1) % 4] * by[k];
    double sgn = orient >= 0 ? 1.0 : -1.0;
    for (int e = 0; e < 4 && n > 0; ++e) {
        double ex = bx[(e + 1) % 4] - bx[e], ey = by[(e + 1) % 4] - by[e];
        int m = 0;
        for (int i = 0; i < n; ++i) {
            int j = (i + 1) % n;
            double di = sgn * (ex * (py[i] - by[e]) - ey * (px[i] - bx[e]));
            double dj = sgn * (ex * (py[j] - by[e]) - ey * (px[j] - bx[e]));
            if (di >= 0) { qx[m] = px[i]; qy[m] = py[i]; ++m; }
            if ((di >= 0) != (dj >= 0)) {
                double t = di / (di - dj);
                qx[m] = px[i] + t * (px[j] - px[i]);
                qy[m] = py[i] + t * (py[j] - py[i]);
                ++m;
            }
        }
        n = m;
        memcpy(px, qx, sizeof(double) * n);
        memcpy(py, qy, sizeof(double) * n);
    }
    double area = 0;
    for (int i = 0; i < n; ++i) area += px[i] * py[(i + 1) % n] - px[(i + 1) % n] * py[i];
    return fabs(area) / 2;
}

ORC_API void orc_boxes_iou_bev64(const float *a, int na, const float *b, int nb, double *out)
{
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) {
            const float *p = a + i * 5, *q = b + j * 5;
            double sa = ((double)p[2] - p[0]) * ((double)p[3] - p[1]);
            double sb = ((double)q[2] - q[0]) * ((double)q[3] - q[1]);
            double so = orc_box_overlap64(p, q);
            double den = sa + sb - so;
            out[(size_t)i * nb + j] = so / (den > 1e-8 ? den : 1e-8);
        }
}

/* ---------------------------------------------------------------------------------------------
 * RoI-aware point pooling: restatement of pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu
 * (check_pt_in_box3d :25-40, generate_pts_mask_for_box3d :43-81, collect_inside_pts_for_box3d :84-114,
 * roiaware_maxpool3d :116-167, roiaware_avgpool3d :170-199, points_in_boxes_kernel :312-333), serial C,
 * same fp32/double mix of the expressions.  pts_idx_of_voxels / pooled / argmax zeroed by the caller.
 * ------------------------------------------------------------------------------------------- */
static int orc_pt_in_box3d(const float *pt, const float *box, float *local_x, float *local_y)
{
    float x = pt[0], y = pt[1], z = pt[2];
    float cx = box[0], cy = box[1], cz = box[2];
    float w = box[3], l = box[4], h = box[5], rz = box[6];
    cz += h / 2.0;
    if (fabsf(z - cz) > h / 2.0) return 0;
    float rot_angle = rz + 3.14159265358979323846 / 2;      /* M_PI */
    float cosa = cosf(rot_angle), sina = sinf(rot_angle);
    float sx = x - cx, sy = y - cy;
    *local_x = sx * cosa + sy * (-sina);
    *local_y = sx * sina + sy * cosa;
    return (*local_x > -l / 2.0) & (*local_x < l / 2.0) & (*local_y > -w / 2.0) & (*local_y < w / 2.0);
}

ORC_API void orc_roiaware_pool3d(const float *rois, int n_rois, const float *pts, int n_pts, const float *feat, int channels,
                         int out_x, int out_y, int out_z, int max_pts, int pool_method, int32_t *argmax,
                         int32_t *pts_idx_of_voxels, float *pooled)
{
    const int n_vox = out_x * out_y * out_z, max_num = max_pts - 1;
    for (int b = 0; b < n_rois; ++b) {
        const float *box = rois + (size_t)b * 7;
        int32_t *lists = pts_idx_of_voxels + (size_t)b * n_vox * max_pts;
        const float w = box[3], l = box[4], h = box[5];
        for (int k = 0; k < n_pts; ++k) {
            float lx = 0, ly = 0;
            if (!orc_pt_in_box3d(pts + (size_t)k * 3, box, &lx, &ly)) continue;
            float lz = pts[(size_t)k * 3 + 2] - box[2];
            float x_res = l / out_x, y_res = w / out_y, z_res = h / out_z;
            unsigned int xi = (int)((lx + l / 2) / x_res), yi = (int)((ly + w / 2) / y_res), zi = (int)(lz / z_res);
            if (xi > (unsigned)(out_x - 1)) xi = out_x - 1;
            if (yi > (unsigned)(out_y - 1)) yi = out_y - 1;
            if (zi > (unsigned)(out_z - 1)) zi = out_z - 1;
            int32_t *list = lists + ((size_t)((xi & 0xFF) * out_y + (yi & 0xFF)) * out_z + (zi & 0xFF)) * max_pts;
            if (list[0] < max_num) { list[list[0] + 1] = k; list[0]++; }
        }
        for (int v = 0; v < n_vox; ++v) {
            const int32_t *list = lists + (size_t)v * max_pts;
            for (int c = 0; c < channels; ++c) {
                const size_t o = ((size_t)b * n_vox + v) * channels + c;
                if (pool_method == 0) {
                    int arg = -1;
                    float best = -INFINITY;
                    for (int k = 1; k <= list[0]; ++k) {
                        float val = feat[(size_t)list[k] * channels + c];
                        if (val > best) { best = val; arg = list[k]; }
                    }
                    if (arg != -1) pooled[o] = best;
                    argmax[o] = arg;
                } else {
                    float sum = 0;
                    for (int k = 1; k <= list[0]; ++k) sum += feat[(size_t)list[k] * channels + c];
                    if (list[0] > 0) pooled[o] = sum / list[0];
                }
            }
        }
    }
}

ORC_API void orc_points_in_boxes(const float *boxes, int batch, int n_boxes, const float *pts, int n_pts, int32_t *box_idx)
{
    for (int b = 0; b < batch; ++b)
        for (int p = 0; p < n_pts; ++p) {
            box_idx[(size_t)b * n_pts + p] = -1;
            for (int k = 0; k < n_boxes; ++k) {
                float lx, ly;
                if (orc_pt_in_box3d(pts + ((size_t)b * n_pts + p) * 3, boxes + ((size_t)b * n_boxes + k) * 7, &lx, &ly)) {
                    box_idx[(size_t)b * n_pts + p] = k;
                    break;
                }
            }
        }
}
