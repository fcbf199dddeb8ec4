/* oracle/sfc_oracle.c -- CPU restatement of the reference's Safe Flight Corridor construction.
 *
 * TEST INFRASTRUCTURE ONLY (see lscqp_oracle.h): the product never includes, links or calls this.
 *
 * What is restated (all paths relative to /root/reference):
 *   map     : MapManager::updateOctreeFromCSV  src/map_manager.cpp:262-305  (world CSV boxes -> occupied 0.1 m cells)
 *             MapManager::setGlobalMap         src/map_manager.cpp:59-80    (DynamicEDTOctomap(maxdist = 1.0, ..., unknown = free))
 *   query   : CollisionConstraints::isObstacleInSFC   src/collision_constraints.cpp:779-809
 *   growth  : CollisionConstraints::expandSFC         :820-881 (fixed axis order), :883-946 (goal-directed order)
 *             CollisionConstraints::setAxisCand       :1134-1170
 *   callers : initializeSFC :366-383, constructSFCFromPoint :396-411 + expandSFCFromPoint :669-694,
 *             constructSFCFromConvexHull :413-436 + expandSFCFromConvexHull :696-733 / :735-777
 *   boxes   : Box::isPointInBox :81-88, include :177-179, intersection :190-197, closestPoint :199-210,
 *             isSuperSetOfConvexHull :135-150
 *
 * PARITY PIN STATUS: UNPINNED.  The distance queries of the reference go through dynamicEDT3D (and the occupancy through
 * octomap); neither library is vendored in /root/reference nor installed here, and the reference ships no SFC fixture.
 * Their published behaviour is restated:
 *   - octomap::OcTree::coordToKey: cell index = floor(coordinate * (1 / resolution)) in double;
 *   - insertPointCloud marks exactly the cells holding an inserted point as occupied (hits win over misses);
 *   - DynamicEDTOctomap::getDistanceAndClosestObstacle returns the centre of the Euclidean-nearest occupied cell when
 *     it is closer than maxdist, and leaves `closest` untouched otherwise (also for a point outside the map) -- the
 *     caller's point3d is default-constructed, so such a query sees a phantom obstacle cell at the world origin.
 *   One thing the published behaviour does not fix: WHICH cell is returned when several occupied cells are equally near
 *   (it depends on the brushfire's queue order).  Here: the lowest (x, then y, then z) index among them.
 * Arithmetic follows the reference's types: box corners and points are float (octomap::point3d), the resolution, the
 * margin and every product with them are double and narrowed to float on assignment.  Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SP_EPSILON_FLOAT 1e-5

typedef struct orc_map {
    int n[3];            /* cells per axis */
    int key0[3];         /* octree key (floor(coord / res)) of cell 0 */
    double res, inv_res; /* inv_res = 1.0 / res (octomap's resolution_factor) */
    int maxd2;           /* (maxdist / res)^2 in cells^2: nearest obstacle data is valid below it */
    unsigned char *occ;  /* [nx][ny][nz] */
    int *closest;        /* [nx][ny][nz][3] nearest occupied cell (indices), or -1 */
    float world_min[3], world_max[3];
} orc_map;

static int key_of(const orc_map *m, float coord) { return (int) floor(m->inv_res * (double) coord); }

void orc_map_free(orc_map *m) {
    if (!m) return;
    free(m->occ); free(m->closest); free(m);
}

/* boxes: [n_boxes][6] = centre xyz, size xyz (one row of a world CSV) */
orc_map *orc_map_build(const double *boxes, int n_boxes, double res, const float *world_min, const float *world_max,
                       double maxdist) {
    orc_map *m = (orc_map *) calloc(1, sizeof(orc_map));
    m->res = res; m->inv_res = 1.0 / res;
    for (int k = 0; k < 3; k++) {
        m->world_min[k] = world_min[k]; m->world_max[k] = world_max[k];
        m->key0[k] = key_of(m, world_min[k]);                       /* DynamicEDTOctomap: bbxMin / bbxMax keys */
        m->n[k] = key_of(m, world_max[k]) - m->key0[k] + 1;
    }
    m->maxd2 = (int) pow(maxdist / res, 2);
    const size_t cells = (size_t) m->n[0] * m->n[1] * m->n[2];
    m->occ = (unsigned char *) calloc(cells, 1);
    m->closest = (int *) malloc(cells * 3 * sizeof(int));
    /* map_manager.cpp:262-305: float centre / size, round() of the double quotient, cells [start, end) */
    for (int b = 0; b < n_boxes; b++) {
        int lo[3], hi[3];
        for (int k = 0; k < 3; k++) {
            const float com = (float) boxes[b * 6 + k], size = (float) boxes[b * 6 + 3 + k];
            lo[k] = (int) round(((double) com - 0.5 * (double) size) / res);
            hi[k] = (int) round(((double) com + 0.5 * (double) size) / res);
        }
        for (int i = lo[0]; i < hi[0]; i++)
            for (int j = lo[1]; j < hi[1]; j++)
                for (int k = lo[2]; k < hi[2]; k++) {
                    /* the inserted point is the cell centre; its key is the cell it lies in */
                    const float p[3] = {(float) ((i + 0.5) * res), (float) ((j + 0.5) * res), (float) ((k + 0.5) * res)};
                    int c[3], in = 1;
                    for (int a = 0; a < 3; a++) { c[a] = key_of(m, p[a]) - m->key0[a]; in = in && c[a] >= 0 && c[a] < m->n[a]; }
                    if (in) m->occ[((size_t) c[0] * m->n[1] + c[1]) * m->n[2] + c[2]] = 1;
                }
    }
    for (size_t c = 0; c < cells; c++) m->closest[c * 3] = -2;     /* computed on first use (nearest_cell) */
    return m;
}

/* Euclidean-nearest occupied cell of cell (x, y, z), valid below maxdist; ties: lowest (x, y, z).  Brute force over the
 * window, cached. */
static const int *nearest_cell(const orc_map *m, int x, int y, int z) {
    int *c = m->closest + (((size_t) x * m->n[1] + y) * m->n[2] + z) * 3;
    if (c[0] != -2) return c;
    const int R = (int) ceil(sqrt((double) m->maxd2));
    int best = m->maxd2, bx = -1, by = -1, bz = -1;
    for (int i = (x - R < 0 ? 0 : x - R); i <= x + R && i < m->n[0]; i++)
        for (int j = (y - R < 0 ? 0 : y - R); j <= y + R && j < m->n[1]; j++)
            for (int k = (z - R < 0 ? 0 : z - R); k <= z + R && k < m->n[2]; k++) {
                if (!m->occ[((size_t) i * m->n[1] + j) * m->n[2] + k]) continue;
                const int d2 = (i - x) * (i - x) + (j - y) * (j - y) + (k - z) * (k - z);
                if (d2 < best) { best = d2; bx = i; by = j; bz = k; }
            }
    c[0] = bx; c[1] = by; c[2] = bz;
    return c;
}

int orc_map_dims(const orc_map *m, int *n3, int *key0) {
    for (int k = 0; k < 3; k++) { n3[k] = m->n[k]; key0[k] = m->key0[k]; }
    return m->maxd2;
}
const unsigned char *orc_map_occupancy(const orc_map *m) { return m->occ; }
/* nearest occupied cell of every cell ([nx][ny][nz][3], -1 = none within maxdist): fills the whole cache */
const int *orc_map_closest(const orc_map *m) {
    for (int x = 0; x < m->n[0]; x++) for (int y = 0; y < m->n[1]; y++) for (int z = 0; z < m->n[2]; z++) nearest_cell(m, x, y, z);
    return m->closest;
}

/* centre of the closest obstacle cell as getDistanceAndClosestObstacle leaves it in the caller's point3d */
static void closest_obstacle(const orc_map *m, const float *p, float *closest) {
    closest[0] = closest[1] = closest[2] = 0.0f;                   /* default-constructed point3d */
    int c[3];
    for (int k = 0; k < 3; k++) {
        c[k] = key_of(m, p[k]) - m->key0[k];
        if (c[k] < 0 || c[k] >= m->n[k]) return;                    /* outside the map: untouched */
    }
    const int *o = nearest_cell(m, c[0], c[1], c[2]);
    if (o[0] < 0) return;                                           /* nothing within maxdist: untouched */
    for (int k = 0; k < 3; k++) closest[k] = (float) (((double) (o[k] + m->key0[k]) + 0.5) * m->res);   /* keyToCoord */
}

/* collision_constraints.cpp:779-809 */
int orc_is_obstacle_in_sfc(const orc_map *m, const float *box /* min xyz, max xyz */, double margin) {
    const float delta = (float) (0.5 * m->res);
    int size[3];
    for (int i = 0; i < 3; i++) {
        const float ext = box[3 + i] - box[i];
        size[i] = (int) floor(((double) ext + SP_EPSILON_FLOAT) / m->res) + 1;
    }
    /* (an inverted box gives a negative count, which the reference's size_t loop variable turns into a huge one:
     *  undefined there, no grid point here) */
    for (int a = 0; a < size[0]; a++)
        for (int b = 0; b < size[1]; b++)
            for (int c = 0; c < size[2]; c++) {
                const int it[3] = {a, b, c};
                float sp[3], cl[3];
                for (int i = 0; i < 3; i++) sp[i] = (float) ((double) box[i] + (double) it[i] * m->res);
                closest_obstacle(m, sp, cl);
                double dist = 0;
                for (int i = 0; i < 3; i++) {
                    const float lo = cl[i] - delta, hi = cl[i] + delta;      /* Box(closest - delta, closest + delta) */
                    float cp = sp[i];
                    if (sp[i] < lo) cp = lo; else if (sp[i] > hi) cp = hi;   /* Box::closestPoint */
                    const float d = cp - sp[i];
                    const double ad = fabs((double) d);                       /* LInfinityDistance */
                    if (dist < ad) dist = ad;
                }
                if (dist < margin + SP_EPSILON_FLOAT) return 1;
            }
    return 0;
}

/* collision_constraints.cpp:811-818 with margin 0 */
static int in_boundary(const orc_map *m, const float *box) {
    for (int k = 0; k < 3; k++) {
        if (!((double) box[k] > (double) m->world_min[k] + 0 - SP_EPSILON_FLOAT)) return 0;
        if (!((double) box[3 + k] < (double) m->world_max[k] - 0 + SP_EPSILON_FLOAT)) return 0;
    }
    return 1;
}

/* collision_constraints.cpp:1134-1170 */
static void set_axis_cand(const float *box, const float *goal, int *axis_cand) {
    int offsets[3], order[3], n = 0;
    double values[3];
    for (int k = 0; k < 3; k++) {
        const float mid = (box[k] + box[3 + k]) * 0.5f;               /* (box_min + box_max) * 0.5 in float */
        const float d = goal[k] - mid;
        offsets[k] = d > 0 ? 3 : 0;
        values[k] = fabs((double) d);                                 /* abs(float) pushed into a vector<double> */
    }
    double max_value = -1, min_value = 1e9;                           /* SP_INFINITY */
    for (int i = 0; i < 3; i++) {
        int pos;
        if (values[i] > max_value) { pos = 0; max_value = values[i]; }
        else if (values[i] < min_value) { pos = n; min_value = values[i]; }
        else pos = 1;
        for (int j = n; j > pos; j--) order[j] = order[j - 1];
        order[pos] = i; n++;
    }
    for (int i = 0; i < 3; i++) {
        axis_cand[i] = order[i] + offsets[order[i]];
        axis_cand[5 - i] = order[i] + (3 - offsets[order[i]]);
    }
}

/* collision_constraints.cpp:820-881 (goal == NULL) and :883-946 (goal-directed axis order) */
int orc_expand_sfc(const orc_map *m, const float *initial, const float *goal, double margin, float *out) {
    if (orc_is_obstacle_in_sfc(m, initial, margin)) return 0;
    int axis_cand[6] = {0, 1, 2, 3, 4, 5}, n_cand = 6;               /* -x, -y, -z, +x, +y, +z */
    if (goal) set_axis_cand(initial, goal, axis_cand);
    float sfc[6], cand[6], upd[6];
    memcpy(sfc, initial, sizeof(sfc));
    int i = -1;
    while (n_cand > 0) {
        memcpy(cand, sfc, sizeof(sfc)); memcpy(upd, sfc, sizeof(sfc));
        while (in_boundary(m, upd) && !orc_is_obstacle_in_sfc(m, upd, margin)) {
            i++;
            if (i >= n_cand) i = 0;
            const int axis = axis_cand[i];
            memcpy(sfc, cand, sizeof(sfc)); memcpy(upd, cand, sizeof(sfc));
            if (axis < 3) {
                upd[3 + axis] = cand[axis];
                cand[axis] = (float) ((double) cand[axis] - m->res);
                upd[axis] = cand[axis];
            } else {
                upd[axis - 3] = cand[axis];
                cand[axis] = (float) ((double) cand[axis] + m->res);
                upd[axis] = cand[axis];
            }
        }
        if (i < 0) i = 0;   /* (start box already outside the world: the reference erases begin() - 1, undefined; not reached by missions) */
        for (int j = i; j + 1 < n_cand; j++) axis_cand[j] = axis_cand[j + 1];     /* erase(begin + i) */
        n_cand--;
        if (i > 0) i--; else i = n_cand - 1;
    }
    const double delta = margin - ((int) (margin / m->res) * m->res);               /* margin compensation */
    for (int k = 0; k < 3; k++) {
        if ((double) sfc[k] > (double) m->world_min[k] + SP_EPSILON_FLOAT) sfc[k] = (float) ((double) sfc[k] - delta);
        if ((double) sfc[3 + k] < (double) m->world_max[k] - SP_EPSILON_FLOAT) sfc[3 + k] = (float) ((double) sfc[3 + k] + delta);
    }
    memcpy(out, sfc, sizeof(sfc));
    return 1;
}

static int point_in_box(const float *box, const float *p) {            /* :81-88 */
    for (int k = 0; k < 3; k++)
        if (!((double) p[k] > (double) box[k] - SP_EPSILON_FLOAT && (double) p[k] < (double) box[3 + k] + SP_EPSILON_FLOAT)) return 0;
    return 1;
}

/* if the grid-aligned start box leaves the previous corridor: intersect and re-align inwards (:680-688, :764-771) */
static void clip_to_prev(const orc_map *m, const float *prev, float *init) {
    if (point_in_box(prev, init) && point_in_box(prev, init + 3)) return;            /* prev.include(initial) */
    for (int k = 0; k < 3; k++) {
        init[k] = init[k] > prev[k] ? init[k] : prev[k];                              /* intersection */
        init[3 + k] = init[3 + k] < prev[3 + k] ? init[3 + k] : prev[3 + k];
    }
    for (int k = 0; k < 3; k++) {
        init[k] = (float) (ceil(((double) init[k] - SP_EPSILON_FLOAT) / m->res) * m->res);
        init[3 + k] = (float) (floor(((double) init[3 + k] + SP_EPSILON_FLOAT) / m->res) * m->res);
    }
}

/* initializeSFC :366-383: the corridor of every segment on the first replan.  Returns 0 where the reference throws. */
int orc_sfc_initialize(const orc_map *m, const float *position, double radius, float *out) {
    float init[6];
    for (int k = 0; k < 3; k++) {
        init[k] = (float) (floor((double) position[k] / m->res) * m->res);
        init[3 + k] = (float) (ceil((double) position[k] / m->res) * m->res);
    }
    return orc_expand_sfc(m, init, NULL, radius, out);
}

/* constructSFCFromPoint :396-411 (+ expandSFCFromPoint :669-694): corridor of the last segment; on failure the
 * previous one is reused (out = prev), return value 0 */
int orc_sfc_from_point(const orc_map *m, const float *point, const float *goal, const float *prev, double radius, float *out) {
    float init[6];
    for (int k = 0; k < 3; k++) {
        init[k] = (float) (floor((double) point[k] / m->res) * m->res);
        init[3 + k] = (float) (ceil((double) point[k] / m->res) * m->res);
    }
    clip_to_prev(m, prev, init);
    if (orc_expand_sfc(m, init, goal, radius, out)) return 1;
    memcpy(out, prev, 6 * sizeof(float));
    return 0;
}

static int superset_of(const float *box, const float *pts, int n) {    /* isSuperSetOfConvexHull :135-150 */
    for (int i = 0; i < 3; i++) {
        float lo = pts[i], hi = pts[i];
        for (int p = 1; p < n; p++) { if (pts[p * 3 + i] < lo) lo = pts[p * 3 + i]; if (pts[p * 3 + i] > hi) hi = pts[p * 3 + i]; }
        if ((double) lo < (double) box[i] - SP_EPSILON_FLOAT || (double) hi > (double) box[3 + i] + SP_EPSILON_FLOAT) return 0;
    }
    return 1;
}

static void hull_bounds(const float *pts, int n, float *box) {
    for (int k = 0; k < 3; k++) { box[k] = pts[k]; box[3 + k] = pts[k]; }
    for (int p = 0; p < n; p++)
        for (int k = 0; k < 3; k++) {
            if (pts[p * 3 + k] < box[k]) box[k] = pts[p * 3 + k];
            if (pts[p * 3 + k] > box[3 + k]) box[3 + k] = pts[p * 3 + k];
        }
}

/* constructSFCFromConvexHull :413-436: hull = {last point of initial_traj, current goal}; first try with the next
 * waypoint added (:696-733, corners rounded to the grid), then the hull alone clipped to the previous corridor
 * (:735-777); on failure reuse the previous corridor.  Returns 2 / 1 / 0 for the branch taken. */
int orc_sfc_from_convex_hull(const orc_map *m, const float *hull /* [2][3] */, const float *next_waypoint,
                             const float *prev, double radius, float *out) {
    float pts[9], init[6];
    memcpy(pts, hull, 6 * sizeof(float)); memcpy(pts + 6, next_waypoint, 3 * sizeof(float));
    hull_bounds(pts, 3, init);
    for (int k = 0; k < 3; k++) {
        init[k] = (float) (round((double) init[k] / m->res) * m->res);
        init[3 + k] = (float) (round((double) init[3 + k] / m->res) * m->res);
    }
    if (orc_expand_sfc(m, init, NULL, radius, out) && superset_of(out, pts, 3)) return 2;
    hull_bounds(pts, 2, init);
    for (int k = 0; k < 3; k++) {
        init[k] = (float) (floor((double) init[k] / m->res) * m->res);
        init[3 + k] = (float) (ceil((double) init[3 + k] / m->res) * m->res);
    }
    clip_to_prev(m, prev, init);
    if (orc_expand_sfc(m, init, NULL, radius, out)) return 1;           /* (:772-776: success even if not a superset) */
    memcpy(out, prev, 6 * sizeof(float));
    return 0;
}
