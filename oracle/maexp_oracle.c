/* CPU oracle: literal C restatement of the reference's gridworld step + observation path.
 *
 * TEST INFRASTRUCTURE ONLY — see maexp_oracle.h.  Every function cites the reference
 * file:line (relative to /root/reference) it follows.  The data layout deliberately stays
 * the REFERENCE's (int8 one-hot matrices, centred agent coordinates, V x V window slices)
 * so that it is an independent check of the CUDA path, which uses bit-packed rows.
 */
#include "maexp_oracle.h"
#include "philox.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define MAXX 64
#define MAXY 64
#define MAXV 9
#define MAXP 8

typedef struct {
    /* core/worlds/entity.py:47-52 */
    int8_t xy[ORC_MAX_AGENTS][2];      /* centred coordinates */
    int8_t action[ORC_MAX_AGENTS][2];
    uint8_t collide_walls[ORC_MAX_AGENTS], collide_agents[ORC_MAX_AGENTS];
    /* core/maps/abstract_map.py:20-24 (per-env mutable grids) */
    int8_t *agents_matrix;      /* [N,X,Y] */
    int8_t *objects_matrix;     /* [T,X,Y] */
    int8_t *destination;        /* [N,X,Y] */
    int8_t *past_global_agents; /* [P,N,X,Y], transition_agent_observation.py:76-94 */
    int32_t init_xys_order[ORC_MAX_AGENTS]; /* default_environment.py:33 (persists across episodes) */
    int32_t current_step;
    int32_t counters[4];        /* objects_generated, objects_completed, agents_collided, walls_collided */
    int32_t pending_reset;
    uint32_t respawn_draws;     /* draw counter inside the current call */
    /* dynamic_simple_map.py:13-18: obstacle cells and homing flags; map.reset() leaves them alone, so they
     * carry over from episode to episode */
    int16_t obstacle[2][2];
    uint8_t homing[2];
} env_t;

struct orc_env {
    orc_cfg cfg;
    int X, Y, V, R, C, agent_ch, object_ch;
    int8_t wall[MAXX][MAXY];
    int8_t *objects_area; /* [T,X,Y] */
    int8_t noise_area[MAXX][MAXY];
    int8_t *los;          /* [X,Y,V,V] 1 visible / 0 hidden  (local_area_mask.py:97) */
    env_t *envs;
    int32_t *hm;          /* orc_out.heatmaps of the running call (NULL: off) */
};

/* offsets into orc_out.heatmaps */
#define HM_AGENTS(o, k, x, y) ((o)->hm[(((k)) * (o)->X + (x)) * (o)->Y + (y)])
#define HM_COMPLETE(o, k, x, y) ((o)->hm[(((o)->cfg.num_agents + (k)) * (o)->X + (x)) * (o)->Y + (y)])
#define HM_OBJECTS(o, t, x, y) ((o)->hm[((2 * (o)->cfg.num_agents + (t)) * (o)->X + (x)) * (o)->Y + (y)])
#define HM_LEFT(o, t, x, y) ((o)->hm[((2 * (o)->cfg.num_agents + (o)->cfg.type_objects + (t)) * (o)->X + (x)) * (o)->Y + (y)])
#define HM_WALL(o, x, y) ((o)->hm[((2 * (o)->cfg.num_agents + 2 * (o)->cfg.type_objects) * (o)->X + (x)) * (o)->Y + (y)])
#define HM_COLLIDE(o, x, y) ((o)->hm[((2 * (o)->cfg.num_agents + 2 * (o)->cfg.type_objects + 1) * (o)->X + (x)) * (o)->Y + (y)])

#define AM(o, e, k, x, y) ((e)->agents_matrix[((k) * (o)->X + (x)) * (o)->Y + (y)])
#define OM(o, e, t, x, y) ((e)->objects_matrix[((t) * (o)->X + (x)) * (o)->Y + (y)])
#define DM(o, e, k, x, y) ((e)->destination[((k) * (o)->X + (x)) * (o)->Y + (y)])
#define OA(o, t, x, y) ((o)->objects_area[((t) * (o)->X + (x)) * (o)->Y + (y)])
#define LOS(o, x, y, a, b) ((o)->los[((((x) * (o)->Y + (y)) * (o)->V + (a)) * (o)->V) + (b)])
#define PG(o, e, p, k, x, y) ((e)->past_global_agents[((((p) * (o)->cfg.num_agents + (k)) * (o)->X + (x)) * (o)->Y) + (y)])

/* numpy-style negative index wrap for wall_matrix[...] lookups in generate_each_mask */
static int wall_at(const orc_env *o, int x, int y) {
    if (x < 0) x += o->X;
    if (y < 0) y += o->Y;
    if (x < 0 || x >= o->X || y < 0 || y >= o->Y) abort(); /* would be IndexError in the reference */
    return o->wall[x][y];
}

/* ------------------------------------------------------------------ maps */

static void set_cells(int8_t (*w)[MAXY], const int *xs, int nx, const int *ys, int ny, int v) {
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) w[xs[i]][ys[j]] = (int8_t)v;
}

static void area_box(orc_env *o, int t, int x0, int x1, int y0, int y1) {
    for (int x = x0; x < x1; ++x)
        for (int y = y0; y < y1; ++y) OA(o, t, x, y) = 1;
}

/* core/maps/abstract_map.py:68-70, four_rooms_map.py:11-42, three_rooms_map.py:13-40,
 * four_rectangle.py:11-39, simple_map.py:10-21 */
static void build_map(orc_env *o) {
    const int X = o->X, Y = o->Y, T = o->cfg.type_objects;
    memset(o->wall, 0, sizeof(o->wall));
    for (int y = 0; y < Y; ++y) o->wall[0][y] = o->wall[X - 1][y] = 1;
    for (int x = 0; x < X; ++x) o->wall[x][0] = o->wall[x][Y - 1] = 1;
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y) o->noise_area[x][y] = 1; /* abstract_map.py:77-78 */

    switch (o->cfg.map_id) {
    case ORC_MAP_DYNAMIC_SIMPLE:
        /* dynamic_simple_map.py:13-27: obstacles at [X/2, 1] and [X/2, Y/2]; DynamicSimpleMap.step() has no
         * caller in the reference, so the obstacles never move.  Areas are SimpleMap's (fall through). */
        o->wall[X / 2][1] = o->wall[X / 2][Y / 2] = 1;
        /* fall through */
    case ORC_MAP_SIMPLE:
        if (T == 1) area_box(o, 0, 1, X - 1, 1, Y - 1);
        else if (T == 2) { area_box(o, 0, 1, X / 2, 1, Y - 1); area_box(o, 1, X / 2, X - 1, 1, Y - 1); }
        break;
    case ORC_MAP_FOUR_ROOMS: {
        static const int seg[] = {1, 2, 3, 7, 8, 9, 10, 13, 14, 15, 16, 20, 21, 22};
        const int c10 = 10, c13 = 13;
        set_cells(o->wall, seg, 14, &c10, 1, 1);
        set_cells(o->wall, seg, 14, &c13, 1, 1);
        set_cells(o->wall, &c10, 1, seg, 14, 1);
        set_cells(o->wall, &c13, 1, seg, 14, 1);
        if (T == 1) {
            area_box(o, 0, 1, 10, 1, 10); area_box(o, 0, 1, 10, 14, Y - 1);
            area_box(o, 0, 14, X - 1, 1, 10); area_box(o, 0, 14, X - 1, 14, Y - 1);
        } else if (T == 2) {
            area_box(o, 0, 1, 10, 1, 10); area_box(o, 0, 1, 10, 14, Y - 1);
            area_box(o, 1, 14, X - 1, 1, 10); area_box(o, 1, 14, X - 1, 14, Y - 1);
        }
        break;
    }
    case ORC_MAP_THREE_ROOMS: {
        static const int a[] = {1, 2, 3, 7, 8, 9, 10};
        static const int b[] = {1, 2, 3, 7, 8, 9, 10, 14, 15, 16, 17, 21, 22, 23};
        static const int c[] = {1, 2, 3, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 21, 22, 23};
        const int c10 = 10, c14 = 14;
        set_cells(o->wall, a, 7, &c10, 1, 1);
        set_cells(o->wall, a, 7, &c14, 1, 1);
        set_cells(o->wall, &c10, 1, b, 14, 1);
        set_cells(o->wall, &c14, 1, c, 17, 1);
        if (T == 1) {
            area_box(o, 0, 1, 10, 1, 10); area_box(o, 0, 1, 10, 15, 24); area_box(o, 0, 15, 24, 1, 24);
        } else if (T == 2) {
            area_box(o, 0, 1, 10, 1, 10); area_box(o, 0, 15, 24, 1, 24);
            area_box(o, 1, 1, 10, 15, 24); area_box(o, 1, 15, 24, 1, 24);
        }
        break;
    }
    case ORC_MAP_FOUR_RECTANGLE: {
        for (int y = 0; y < Y; ++y) o->wall[19][y] = o->wall[20][y] = 1;
        o->wall[19][11] = o->wall[19][12] = o->wall[20][11] = o->wall[20][12] = 0;
        for (int x = 0; x < X; ++x) o->wall[x][10] = o->wall[x][13] = 1;
        for (int x = 8; x < 12; ++x) o->wall[x][10] = o->wall[x][13] = 0;
        for (int x = 28; x < 32; ++x) o->wall[x][10] = o->wall[x][13] = 0;
        if (T == 1) {
            area_box(o, 0, 1, 19, 1, 10); area_box(o, 0, 1, 19, 14, Y - 1);
            area_box(o, 0, 21, X - 1, 1, 10); area_box(o, 0, 21, X - 1, 14, Y - 1);
        } else if (T == 2) {
            area_box(o, 0, 1, 19, 1, 10); area_box(o, 0, 1, 19, 14, Y - 1);
            area_box(o, 1, 21, X - 1, 1, 10); area_box(o, 1, 21, X - 1, 14, Y - 1);
        }
        break;
    }
    /* central_room_map.py:10-52 (every type_objects branch only fills type 0) and central_room_destination_map.py */
    case ORC_MAP_CENTRAL_ROOM:
    case ORC_MAP_CENTRAL_ROOM_DESTINATION: {
        for (int v = 8; v < 17; ++v) { o->wall[v][8] = o->wall[v][16] = 1; o->wall[8][v] = o->wall[16][v] = 1; }
        for (int v = 11; v < 14; ++v) { o->wall[v][8] = o->wall[v][16] = 0; o->wall[8][v] = o->wall[16][v] = 0; }
        if (T >= 1 && T <= 5) {
            area_box(o, 0, 1, 24, 1, 24);
            for (int x = 8; x < 17; ++x)
                for (int y = 8; y < 17; ++y) OA(o, 0, x, y) = 0;
        }
        break;
    }
    /* central_room_large_map.py:10-23 */
    case ORC_MAP_CENTRAL_ROOM_LARGE:
    case ORC_MAP_CENTRAL_ROOM_LARGE_DESTINATION: {
        for (int x = 20; x < 29; ++x) o->wall[x][8] = o->wall[x][16] = 1;
        for (int y = 8; y < 17; ++y) o->wall[20][y] = o->wall[28][y] = 1;
        for (int x = 23; x < 26; ++x) o->wall[x][8] = o->wall[x][16] = 0;
        for (int y = 11; y < 14; ++y) o->wall[20][y] = o->wall[28][y] = 0;
        area_box(o, 0, 1, 48, 1, 24);
        for (int x = 20; x < 29; ++x)
            for (int y = 8; y < 17; ++y) OA(o, 0, x, y) = 0;
        break;
    }
    /* two_central_rooms_map.py:10-46 */
    case ORC_MAP_TWO_CENTRAL_ROOMS: {
        for (int v = 8; v < 17; ++v) { o->wall[v][8] = o->wall[v][16] = 1; o->wall[8][v] = o->wall[16][v] = 1; }
        for (int v = 11; v < 14; ++v) { o->wall[v][8] = o->wall[v][16] = 0; o->wall[8][v] = o->wall[16][v] = 0; }
        for (int x = 33; x < 42; ++x) o->wall[x][8] = o->wall[x][16] = 1;
        for (int y = 8; y < 17; ++y) o->wall[33][y] = o->wall[41][y] = 1;
        for (int x = 36; x < 39; ++x) o->wall[x][8] = o->wall[x][16] = 0;
        for (int y = 11; y < 14; ++y) o->wall[33][y] = o->wall[41][y] = 0;
        if (T == 1) {
            area_box(o, 0, 1, 24, 1, 24);
            for (int x = 8; x < 17; ++x)
                for (int y = 8; y < 17; ++y) OA(o, 0, x, y) = 0;
        } else if (T == 2) {
            area_box(o, 0, 1, 26, 1, 13);
            for (int x = 8; x < 17; ++x)
                for (int y = 8; y < 13; ++y) OA(o, 0, x, y) = 0;
            area_box(o, 0, 24, 49, 12, 24);
            for (int x = 33; x < 42; ++x)
                for (int y = 12; y < 17; ++y) OA(o, 0, x, y) = 0;
            area_box(o, 1, 1, 26, 12, 24);
            for (int x = 8; x < 17; ++x)
                for (int y = 12; y < 17; ++y) OA(o, 1, x, y) = 0;
            area_box(o, 1, 24, 49, 1, 13);
            for (int x = 33; x < 42; ++x)
                for (int y = 8; y < 13; ++y) OA(o, 1, x, y) = 0;
        }
        break;
    }
    default: abort();
    }
}

/* ------------------------------------------------------ line-of-sight table */

/* masks/local_area_mask.py:100-332 (generate_each_mask).  Offsets (cx, cy) are the
 * reference's centred offsets; coord2ind (abstract_map.py:44-53) maps them to
 * window index (R+cx, R-cy) and map index (x+cx, y-cy). */
static void each_mask(const orc_env *o, int x, int y, int8_t m[MAXV][MAXV]) {
    const int R = o->R, V = o->V, X = o->X, Y = o->Y;
    for (int a = 0; a < V; ++a)
        for (int b = 0; b < V; ++b) m[a][b] = -1;
    m[R][R] = 0;
#define OUTSIDE(mx, my) ((mx) < 0 || X <= (mx) || (my) < 0 || Y <= (my))
    /* :107-146 */
    for (int dx = -1; dx < 2; ++dx)
        for (int dyi = 0; dyi < 2; ++dyi)
            for (int opi = 0; opi < 2; ++opi) {
                const int dy = dyi ? 1 : -1, opr = opi ? 1 : -1;
                for (int j = 0; j < R; ++j) {
                    const int cx = dx + j * opr, cy = dy + j * dy;
                    const int lx = R + cx, ly = R - cy, mx = x + cx, my = y - cy;
                    if (OUTSIDE(mx, my)) { m[lx][ly] = -1; break; }
                    if (j == 0 && dx == 0 && wall_at(o, mx + opr, my) == 0) { m[lx][ly] = 0; continue; }
                    if (j == 0 && dx != 0 && opr != dx) {
                        if (wall_at(o, mx - dx, my) == 0) continue;
                        else if (wall_at(o, mx - dx, my) == 1) break;
                    }
                    if (wall_at(o, mx, my) == 1) { m[lx][ly] = -1; break; }
                    m[lx][ly] = 0;
                }
            }
    /* :150-191 */
    for (int dy = -1; dy < 2; ++dy)
        for (int dxi = 0; dxi < 2; ++dxi)
            for (int opi = 0; opi < 2; ++opi) {
                const int dx = dxi ? 1 : -1, opr = opi ? 1 : -1;
                for (int j = 0; j < R; ++j) {
                    const int cx = dx + j * dx, cy = dy + j * opr;
                    const int lx = R + cx, ly = R - cy, mx = x + cx, my = y - cy;
                    if (OUTSIDE(mx, my)) { m[lx][ly] = -1; break; }
                    if (j == 0 && dy == 0 && wall_at(o, mx, my - opr) == 0) { m[lx][ly] = 0; continue; }
                    if (j == 0 && dy != 0 && opr != dy) {
                        if (wall_at(o, mx, my + dy) == 0) continue;
                        else if (wall_at(o, mx, my + dy) == 1) break;
                    }
                    if (wall_at(o, mx, my) == 1) { m[lx][ly] = -1; break; }
                    m[lx][ly] = 0;
                }
            }
    /* :193-261 */
    for (int oxi = 0; oxi < 2; ++oxi) {
        const int opr_x = oxi ? 1 : -1;
        for (int dx = 0; dx < R + 1; ++dx) {
            {
                const int cx = dx * opr_x;
                if (wall_at(o, x + cx, y) == 1) { m[R + cx][R] = -1; break; }
            }
            for (int oyi = 0; oyi < 2; ++oyi) {
                const int opr_y = oyi ? 1 : -1;
                for (int dy = 0; dy < 2; ++dy) {
                    const int cx = dx * opr_x, cy = dy * opr_y;
                    const int lx = R + cx, ly = R - cy, mx = x + cx, my = y - cy;
                    if (OUTSIDE(mx, my)) { m[lx][ly] = -1; continue; }
                    if (wall_at(o, mx, my) == 1) { m[lx][ly] = -1; break; }
                    else m[lx][ly] = 0;
                }
            }
        }
        for (int oyi = 0; oyi < 2; ++oyi) {
            const int opr_y = oyi ? 1 : -1;
            for (int dx = 0; dx < R + 1; ++dx) {
                const int cx = dx * opr_x, cy = opr_y;
                const int lx = R + cx, ly = R - cy, mx = x + cx, my = y - cy;
                if (OUTSIDE(mx, my)) { m[lx][ly] = -1; continue; }
                if (wall_at(o, mx, my) == 1) { m[lx][ly] = -1; break; }
                else m[lx][ly] = 0;
            }
        }
    }
    /* :263-330 */
    for (int oyi = 0; oyi < 2; ++oyi) {
        const int opr_y = oyi ? 1 : -1;
        for (int dy = 0; dy < R + 1; ++dy) {
            {
                const int cy = dy * opr_y;
                if (wall_at(o, x, y - cy) == 1) { m[R][R - cy] = -1; break; }
            }
            for (int oxi = 0; oxi < 2; ++oxi) {
                const int opr_x = oxi ? 1 : -1;
                for (int dx = 0; dx < 2; ++dx) {
                    const int cx = dx * opr_x, cy = dy * opr_y;
                    const int lx = R + cx, ly = R - cy, mx = x + cx, my = y - cy;
                    if (OUTSIDE(mx, my)) { m[lx][ly] = -1; continue; }
                    if (wall_at(o, mx, my) == 1) { m[lx][ly] = -1; break; }
                    else m[lx][ly] = 0;
                }
            }
        }
        for (int oxi = 0; oxi < 2; ++oxi) {
            const int opr_x = oxi ? 1 : -1;
            for (int dy = 0; dy < R + 1; ++dy) {
                const int cx = opr_x, cy = dy * opr_y;
                const int lx = R + cx, ly = R - cy, mx = x + cx, my = y - cy;
                if (OUTSIDE(mx, my)) { m[lx][ly] = -1; continue; }
                if (wall_at(o, mx, my) == 1) { m[lx][ly] = -1; break; }
                else m[lx][ly] = 0;
            }
        }
    }
#undef OUTSIDE
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* masks/local_area_mask.py:18-47 for an index-space position */
static void mask_coordinates(const orc_env *o, int px, int py, int32_t c[12]) {
    const int R = o->R, V = o->V;
    c[0] = abs(imin(0, px - R));
    c[1] = V - imax(0, px + R - (o->X - 1));
    c[2] = abs(imin(0, py - R));
    c[3] = V - imax(0, py + R - (o->Y - 1));
    c[4] = c[0]; c[5] = c[1]; c[6] = c[2]; c[7] = c[3];
    c[8] = imax(0, px - R);
    c[9] = imin(o->X - 1, px + R) + 1;
    c[10] = imax(0, py - R);
    c[11] = imin(o->Y - 1, py + R) + 1;
}

/* masks/local_area_mask.py:50-97 */
static void build_los(orc_env *o) {
    const int X = o->X, Y = o->Y, V = o->V;
    for (size_t i = 0; i < (size_t)X * Y * V * V; ++i) o->los[i] = -1;
    for (int x = 1; x < X - 1; ++x)
        for (int y = 1; y < Y - 1; ++y) {
            if (o->cfg.transparent_observation) {
                int32_t c[12];
                mask_coordinates(o, x, y, c);
                for (int a = c[0]; a < c[1]; ++a)
                    for (int b = c[2]; b < c[3]; ++b)
                        LOS(o, x, y, a, b) = (int8_t)(o->wall[c[8] + a - c[0]][c[10] + b - c[2]] * -1);
            } else {
                int8_t m[MAXV][MAXV];
                each_mask(o, x, y, m);
                for (int a = 0; a < V; ++a)
                    for (int b = 0; b < V; ++b) LOS(o, x, y, a, b) = m[a][b];
            }
        }
    for (size_t i = 0; i < (size_t)X * Y * V * V; ++i) o->los[i] += 1;
}

/* wall_matrix as this environment sees it NOW.  o->wall is the matrix at construction time (the line-of-sight
 * table is generated from it once, local_area_mask.py:18-97); on dynamic_simple the two obstacle cells move
 * with every world step (dynamic_simple_map.py:29-63) while the border stays. */
static int env_wall(const orc_env *o, const env_t *en, int x, int y) {
    if (o->cfg.map_id != ORC_MAP_DYNAMIC_SIMPLE) return o->wall[x][y];
    if (x == 0 || y == 0 || x == o->X - 1 || y == o->Y - 1) return 1;
    for (int k = 0; k < 2; ++k)
        if (en->obstacle[k][0] == x && en->obstacle[k][1] == y) return 1;
    return 0;
}

/* dynamic_simple_map.py:29-63, called first thing in world.step (abstract_world.py:40) */
static void map_step(const orc_env *o, env_t *en) {
    if (o->cfg.map_id != ORC_MAP_DYNAMIC_SIMPLE) return;
    const int X = o->X, Y = o->Y;
    for (int k = 0; k < 2; ++k) {
        const int x = en->obstacle[k][0], y = en->obstacle[k][1];
        if ((x == X / 2 && y == Y - 2 && k == 0) || (x == X - 2 && y == Y / 2 && k == 1)) en->homing[k] = 1;
        else if ((x == X / 2 && y == 1 && k == 0) || (x == 1 && y == Y / 2 && k == 1)) en->homing[k] = 0;
    }
    en->obstacle[0][1] = (int16_t)(en->obstacle[0][1] + (en->homing[0] ? -1 : 1));
    en->obstacle[1][0] = (int16_t)(en->obstacle[1][0] + (en->homing[1] ? -1 : 1));
}

/* ------------------------------------------------------------ construction */

orc_env *orc_create(const orc_cfg *cfg) {
    orc_env *o = (orc_env *)calloc(1, sizeof(orc_env));
    o->cfg = *cfg;
    switch (cfg->map_id) { /* core/maps/__init__.py:26-39 */
    case ORC_MAP_DYNAMIC_SIMPLE:
    case ORC_MAP_SIMPLE: o->X = cfg->size_x; o->Y = cfg->size_y; break;
    case ORC_MAP_FOUR_ROOMS: o->X = 24; o->Y = 24; break;
    case ORC_MAP_THREE_ROOMS: o->X = 25; o->Y = 25; break;
    case ORC_MAP_FOUR_RECTANGLE: o->X = 40; o->Y = 24; break;
    case ORC_MAP_CENTRAL_ROOM: case ORC_MAP_CENTRAL_ROOM_DESTINATION: o->X = 25; o->Y = 25; break;
    case ORC_MAP_CENTRAL_ROOM_LARGE: case ORC_MAP_CENTRAL_ROOM_LARGE_DESTINATION: o->X = 49; o->Y = 25; break;
    case ORC_MAP_TWO_CENTRAL_ROOMS: o->X = 50; o->Y = 25; break;
    default: free(o); return NULL;
    }
    o->V = cfg->visible_range;
    o->R = o->V / 2;
    const int N = cfg->num_agents, T = cfg->type_objects, P = cfg->past_step;
    switch (cfg->agent_view) { /* agents/*.py get_channel */
    case ORC_AV_DEFAULT: case ORC_AV_SIMPLE: o->agent_ch = 1; break;
    case ORC_AV_INDIVIDUAL: o->agent_ch = N; break;
    case ORC_AV_TRANSITION: o->agent_ch = 1 + P; break;
    case ORC_AV_INDIVIDUAL_TRANSITION: o->agent_ch = N + 1 + P; break;
    default: free(o); return NULL;
    }
    o->object_ch = cfg->object_view == ORC_OV_TYPES ? T : 1;
    o->C = o->agent_ch + o->object_ch + 1; /* observation_handler.py:56-61 */
    const size_t XY = (size_t)o->X * o->Y;
    o->objects_area = (int8_t *)calloc(T * XY, 1);
    o->los = (int8_t *)calloc(XY * o->V * o->V, 1);
    build_map(o);
    build_los(o);
    o->envs = (env_t *)calloc(cfg->num_envs, sizeof(env_t));
    for (int e = 0; e < cfg->num_envs; ++e) {
        env_t *en = &o->envs[e];
        en->agents_matrix = (int8_t *)calloc(N * XY, 1);
        en->objects_matrix = (int8_t *)calloc(T * XY, 1);
        en->destination = (int8_t *)malloc(N * XY);
        memset(en->destination, 1, N * XY); /* abstract_map.py:24 */
        en->past_global_agents = (int8_t *)calloc((size_t)(P > 0 ? P : 1) * N * XY, 1);
        for (int i = 0; i < ORC_MAX_AGENTS; ++i) en->init_xys_order[i] = i; /* default_environment.py:33 */
        en->obstacle[0][0] = (int16_t)(o->X / 2); en->obstacle[0][1] = 1; /* dynamic_simple_map.py:13-17 */
        en->obstacle[1][0] = (int16_t)(o->X / 2); en->obstacle[1][1] = (int16_t)(o->Y / 2);
    }
    return o;
}

void orc_destroy(orc_env *o) {
    if (!o) return;
    for (int e = 0; e < o->cfg.num_envs; ++e) {
        free(o->envs[e].agents_matrix); free(o->envs[e].objects_matrix);
        free(o->envs[e].destination); free(o->envs[e].past_global_agents);
    }
    free(o->envs); free(o->objects_area); free(o->los); free(o);
}

int orc_channels(const orc_env *o) { return o->C; }
int orc_obs_elems(const orc_env *o) { return o->C * o->V * o->V; }
int orc_size_x(const orc_env *o) { return o->X; }
int orc_size_y(const orc_env *o) { return o->Y; }

void orc_get_tables(const orc_env *o, int8_t *wall, int8_t *objects_area, int8_t *los) {
    if (wall)
        for (int x = 0; x < o->X; ++x)
            for (int y = 0; y < o->Y; ++y) wall[x * o->Y + y] = o->wall[x][y];
    if (objects_area) memcpy(objects_area, o->objects_area, (size_t)o->cfg.type_objects * o->X * o->Y);
    if (los) memcpy(los, o->los, (size_t)o->X * o->Y * o->V * o->V);
}

void orc_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out4) {
    philox4x32_10(k0, k1, c0, c1, c2, c3, out4);
}

/* --------------------------------------------------------------- helpers */

/* abstract_map.py:44-53 */
static void coord2ind(const orc_env *o, const int8_t xy[2], int *ix, int *iy) {
    *ix = (o->X / 2) + xy[0];
    *iy = (o->Y / 2) - xy[1];
}

static uint32_t genv(const orc_env *o, int e) { return o->cfg.first_global_env_id + (uint32_t)e; }

/* Fisher-Yates with stream draws; this is what the tape's shuffle() does for the live reference */
static void tape_shuffle(const orc_env *o, int e, uint32_t tick, uint32_t stream, uint32_t *d, int32_t *a, int n) {
    for (int k = n - 1; k >= 1; --k) {
        uint32_t j = maexp_scale(maexp_draw(o->cfg.seed, genv(o, e), tick, stream, (*d)++, 0), (uint32_t)(k + 1));
        int32_t t = a[k]; a[k] = a[j]; a[j] = t;
    }
}

/* default_environment.py:134-148 — one accepted object of `object_type`; two uniforms per attempt */
static void generate_one_object(orc_env *o, int e, uint32_t tick, uint32_t stream, uint32_t *attempt, int object_type) {
    env_t *en = &o->envs[e];
    const int N = o->cfg.num_agents, T = o->cfg.type_objects;
    for (;;) {
        uint32_t w[4];
        philox4x32_10(o->cfg.seed, genv(o, e), tick, stream, (*attempt) >> 1, 0, w);
        const uint32_t u1 = w[((*attempt) & 1) * 2], u2 = w[((*attempt) & 1) * 2 + 1];
        ++(*attempt);
        const int x = 1 + (int)maexp_scale(u1, (uint32_t)(o->X - 1));
        const int y = 1 + (int)maexp_scale(u2, (uint32_t)(o->Y - 1));
        int agents = 0, objects = 0;
        for (int k = 0; k < N; ++k) agents += AM(o, en, k, x, y);
        for (int t = 0; t < T; ++t) objects += OM(o, en, t, x, y);
        if (env_wall(o, en, x, y) == 0 && agents == 0 && objects == 0 && OA(o, object_type, x, y) == 1) {
            OM(o, en, object_type, x, y) = 1;
            if (o->hm) HM_OBJECTS(o, object_type, x, y) += 1; /* :145 */
            en->counters[0] += 1;
            return;
        }
    }
}

/* Candidate list entry m of the maps that redraw destinations at reset, value at (x, y):
 * three_rooms_map.py:43-51, central_room_destination_map.py:14-18, central_room_large_destination_map.py:14-18 */
static int dest_candidate(const orc_env *o, int m, int x, int y) {
    const int X = o->X, Y = o->Y;
    switch (o->cfg.map_id) {
    case ORC_MAP_THREE_ROOMS: { /* note: both axes are split at SIZE_X // 2 */
        const int h = X / 2;
        if ((m == 0 || m == 1) && x < h) return 0;
        if ((m == 2 || m == 3) && x >= h) return 0;
        if ((m == 4 || m == 5) && y < h) return 0;
        if ((m == 6 || m == 7) && y >= h) return 0;
        return 1;
    }
    case ORC_MAP_CENTRAL_ROOM_DESTINATION:
        if (m == 0 && x < X / 2) return 0;
        if (m == 1 && x >= X / 2) return 0;
        if (m == 2 && y < Y / 2) return 0;
        if (m == 3 && y >= Y / 2) return 0;
        return 1;
    case ORC_MAP_CENTRAL_ROOM_LARGE_DESTINATION:
        if (m == 0) return x < X / 2;
        if (m == 1) return x >= X / 2;
        if (m == 2) return y < Y / 2;
        return y >= Y / 2;
    }
    return 1;
}

/* two_central_rooms_map.py:48-69: fixed areas for agent pairs; agents past the twelfth keep zeros */
static int two_central_rooms_destination(int k, int x, int y) {
    const int in_a = (x >= 1 && x < 26 && y >= 1 && y < 13) || (x >= 24 && x < 49 && y >= 12 && y < 24);
    const int in_b = (x >= 1 && x < 26 && y >= 12 && y < 24) || (x >= 24 && x < 49 && y >= 1 && y < 13);
    switch (k / 2) {
    case 0: return in_a;
    case 1: return in_b;
    case 2: return x >= 1 && x < 49 && y >= 1 && y < 13;
    case 3: return x >= 1 && x < 49 && y >= 12 && y < 24;
    case 4: return x >= 1 && x < 26 && y >= 1 && y < 24;
    case 5: return x >= 24 && x < 49 && y >= 1 && y < 24;
    }
    return 0;
}

/* abstract_map.py:80-81 (all ones) and the per-map overrides */
static void reset_destination_area(orc_env *o, int e, uint32_t tick) {
    env_t *en = &o->envs[e];
    const int N = o->cfg.num_agents, X = o->X, Y = o->Y;
    memset(en->destination, 1, (size_t)N * X * Y);
    if (o->cfg.map_id == ORC_MAP_TWO_CENTRAL_ROOMS) {
        for (int k = 0; k < N; ++k)
            for (int x = 0; x < X; ++x)
                for (int y = 0; y < Y; ++y) DM(o, en, k, x, y) = (int8_t)two_central_rooms_destination(k, x, y);
        return;
    }
    int M = 0;
    if (o->cfg.map_id == ORC_MAP_THREE_ROOMS) M = 9;
    else if (o->cfg.map_id == ORC_MAP_CENTRAL_ROOM_DESTINATION) M = 5;
    else if (o->cfg.map_id == ORC_MAP_CENTRAL_ROOM_LARGE_DESTINATION) M = 4;
    if (!M) return;
    int32_t perm[9];
    for (int i = 0; i < M; ++i) perm[i] = i;
    uint32_t d = 0;
    tape_shuffle(o, e, tick, MAEXP_STREAM_DEST, &d, perm, M); /* random.shuffle(destination_area) */
    for (int k = 0; k < N; ++k) {                             /* random.choice(destination_area) per agent */
        const uint32_t pick = maexp_scale(maexp_draw(o->cfg.seed, genv(o, e), tick, MAEXP_STREAM_DEST, d++, 0), (uint32_t)M);
        const int m = perm[pick];
        for (int x = 0; x < X; ++x)
            for (int y = 0; y < Y; ++y) DM(o, en, k, x, y) = (int8_t)dest_candidate(o, m, x, y);
    }
}

/* transition_agent_observation.py:81-94 */
static void get_past_global_agents(orc_env *o, env_t *en, int slot) {
    const int N = o->cfg.num_agents;
    for (int i = 0; i < N; ++i) {
        int px, py;
        int32_t c[12];
        coord2ind(o, en->xy[i], &px, &py);
        mask_coordinates(o, px, py, c);
        for (int x = 0; x < o->X; ++x)
            for (int y = 0; y < o->Y; ++y) {
                int s = 0;
                for (int k = 0; k < N; ++k) s += AM(o, en, k, x, y);
                const int inside = (x >= c[8] && x < c[9] && y >= c[10] && y < c[11]);
                PG(o, en, slot, i, x, y) = (int8_t)(s * inside);
            }
    }
}

static int uses_transition(const orc_env *o) {
    return o->cfg.agent_view == ORC_AV_TRANSITION || o->cfg.agent_view == ORC_AV_INDIVIDUAL_TRANSITION;
}

static float u32_to_unit_float(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

/* one noise draw for observation element k of agent i */
static float noise_draw(const orc_env *o, int e, uint32_t tick, int i, int k, const float *tape, int gaussian) {
    if (tape) return tape[((size_t)e * o->cfg.num_agents + i) * (size_t)orc_obs_elems(o) + k];
    uint32_t w[4];
    if (!gaussian) {
        /* uniform draws (flip noise): 16 bits each, EIGHT per call -- element k: call k >> 3, word (k & 7) >> 1, the low half
         * for even k; u = h / 65536 exactly (DESIGN.md RNG specification; csrc/env_kernel.cuh emit_pair) */
        philox4x32_10(o->cfg.seed, genv(o, e), tick, MAEXP_STREAM_NOISE, (uint32_t)(k >> 3), (uint32_t)i, w);
        const uint32_t word = w[(k & 7) >> 1], h = (k & 1) ? word >> 16 : word & 0xFFFFu;
        return (float)h * (1.0f / 65536.0f);
    }
    philox4x32_10(o->cfg.seed, genv(o, e), tick, MAEXP_STREAM_NOISE, (uint32_t)(k >> 2), (uint32_t)i, w);
    /* Box-Muller on the pair (w[2p], w[2p+1]); bit-exact parity is only claimed for taped draws */
    const int p = (k & 3) >> 1;
    const float u1 = ((float)(w[2 * p] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = u32_to_unit_float(w[2 * p + 1]);
    const float rad = sqrtf(-2.0f * logf(u1));
    const float ang = 6.28318530717958647692f * u2;
    return (k & 1) ? rad * sinf(ang) : rad * cosf(ang);
}

/* noises/*.py add_noise restated per element: only the draw of ring(a,b) survives
 * (flip_observation_noise.py:35-41, dist_observation_noise.py:31-38,55-73). */
static void add_noise(const orc_env *o, int e, uint32_t tick, int i, float *obs, const float *tape) {
    const int V = o->V, C = o->C, kind = o->cfg.noise;
    volatile float tmp; /* keep mul and add separately rounded (no FMA contraction) */
    for (int c = 0; c < C; ++c)
        for (int a = 0; a < V; ++a)
            for (int b = 0; b < V; ++b) {
                const int k = (c * V + a) * V + b;
                int ring = imin(imin(a, b), imin(V - 1 - a, V - 1 - b));
                ring = imin(ring, V / 2 - 1);
                float x = obs[k];
                if (kind == ORC_NOISE_FLIP) {
                    const float u = noise_draw(o, e, tick, i, k, tape, 0);
                    if (c == 2) x = x + 1.0f;                /* flip_observation_noise.py:33 */
                    tmp = u * o->cfg.flip_noise_range[ring];
                    x = x + tmp;
                    x = fabsf(rintf(x));                     /* :43 torch.round = half-to-even */
                    if (c == 2) x = x - 1.0f;                /* :44 */
                } else if (kind == ORC_NOISE_DIST) {
                    const float z = noise_draw(o, e, tick, i, k, tape, 1);
                    tmp = z * o->cfg.std_distribution[ring];
                    x = x + tmp;
                } else if (kind == ORC_NOISE_THRESHOLD_DIST) {
                    const float z = noise_draw(o, e, tick, i, k, tape, 1);
                    tmp = z * o->cfg.std_distribution[ring];
                    const float g = tmp, th = o->cfg.noise_threshold[ring];
                    float n = 0.0f;
                    if (g < -th) n = g;
                    if (th < g) n = g;
                    x = x + n;
                } else if (kind == ORC_NOISE_FLAT) {
                    if (c < 3) { /* flat_observation_noise.py:30 hard-codes 3 channels; C must be 3 */
                        const float z = noise_draw(o, e, tick, i, k, tape, 1);
                        tmp = z * 0.03f;
                        x = x + tmp;
                    }
                }
                obs[k] = x;
            }
}

/* observation_handler.py:67-105 for agent i of env e; `noisy` selects fill_obs_noise (:110-117) */
static void observation_ind(orc_env *o, int e, uint32_t tick, int i, float *obs, int noisy, const float *tape,
                            int32_t *coords_out) {
    env_t *en = &o->envs[e];
    const int N = o->cfg.num_agents, T = o->cfg.type_objects, V = o->V, P = o->cfg.past_step;
    int px, py;
    int32_t c[12];
    coord2ind(o, en->xy[i], &px, &py);
    mask_coordinates(o, px, py, c);
    if (coords_out) memcpy(coords_out, c, sizeof(c));
    const int VV = V * V;
    memset(obs, 0, sizeof(float) * (size_t)o->C * VV);
    int ch = 0;
    const int av = o->cfg.agent_view, ov = o->cfg.object_view;
    /* the slices obs[obs_x_min:obs_x_max, obs_y_min:obs_y_max] <- f(area_mask[...], map[...]) */
#define FOR_WINDOW for (int a = c[0]; a < c[1]; ++a) for (int b = c[2]; b < c[3]; ++b)
#define GX (c[8] + a - c[0])
#define GY (c[10] + b - c[2])
    if (av == ORC_AV_INDIVIDUAL || av == ORC_AV_INDIVIDUAL_TRANSITION) {
        /* individual_agent_observation.py:21-39 */
        for (int k = 0; k < N; ++k) {
            FOR_WINDOW obs[(ch * V + a) * V + b] = (float)(LOS(o, px, py, a, b) * AM(o, en, k, GX, GY));
            ++ch;
        }
    }
    if (av == ORC_AV_DEFAULT || av == ORC_AV_TRANSITION || av == ORC_AV_INDIVIDUAL_TRANSITION) {
        /* default_agent_observation.py:21-39, transition_agent_observation.py:29-45 */
        FOR_WINDOW {
            int s = 0;
            for (int k = 0; k < N; ++k) s += AM(o, en, k, GX, GY);
            obs[(ch * V + a) * V + b] = (float)(LOS(o, px, py, a, b) * s);
        }
        ++ch;
    }
    if (av == ORC_AV_SIMPLE) {
        /* simple_agent_observation.py:22-36 */
        FOR_WINDOW {
            int s = 0;
            for (int k = 0; k < N; ++k) s += AM(o, en, k, GX, GY);
            obs[(ch * V + a) * V + b] = (float)s;
        }
        ++ch;
    }
    if (av == ORC_AV_TRANSITION || av == ORC_AV_INDIVIDUAL_TRANSITION) {
        /* transition_agent_observation.py:47-57 */
        for (int t = 0; t < P; ++t) {
            FOR_WINDOW obs[(ch * V + a) * V + b] += (float)PG(o, en, t, i, GX, GY);
            ++ch;
        }
    }
    if (ov == ORC_OV_DEFAULT || ov == ORC_OV_SIMPLE) {
        /* objects/default_object_observation.py:20-39, simple_object_observation.py:21-36 */
        FOR_WINDOW {
            int s = 0;
            for (int t = 0; t < T; ++t) s += OM(o, en, t, GX, GY);
            obs[(ch * V + a) * V + b] = (float)((ov == ORC_OV_DEFAULT ? LOS(o, px, py, a, b) : 1) * s);
        }
        ++ch;
    } else {
        /* objects/types_object_observation.py:20-38 */
        for (int t = 0; t < T; ++t) {
            FOR_WINDOW obs[(ch * V + a) * V + b] = (float)(LOS(o, px, py, a, b) * OM(o, en, t, GX, GY));
            ++ch;
        }
    }
    /* observation_handler.py:107-108 */
    for (int a = 0; a < V; ++a)
        for (int b = 0; b < V; ++b) obs[(ch * V + a) * V + b] = (float)LOS(o, px, py, a, b) - 1.0f;
#undef FOR_WINDOW
#undef GX
#undef GY
    if (noisy && o->cfg.noise != ORC_NOISE_NONE && o->noise_area[px][py] != 0) add_noise(o, e, tick, i, obs, tape);
}

static void emit_observation(orc_env *o, int e, uint32_t tick, int i, const orc_out *out) {
    if (!out) return;
    env_t *en = &o->envs[e];
    const int N = o->cfg.num_agents, T = o->cfg.type_objects;
    const size_t K = (size_t)orc_obs_elems(o), XY = (size_t)o->X * o->Y;
    const size_t slot = (size_t)e * N + i;
    float scratch[40 * MAXV * MAXV];
    int32_t cbuf[12];
    float *dst = out->obs ? out->obs + slot * K : scratch;
    /* merged: the local handler carries the noise, the relative one never does
     * (merged_observtion_handler.py:21-31,41-47); relative alone: same crop, no noise support */
    const int noisy = (o->cfg.area_mask != ORC_MASK_RELATIVE);
    observation_ind(o, e, tick, i, dst, noisy, out->noise_tape, cbuf);
    if (o->cfg.area_mask == ORC_MASK_MERGED && out->obs_rel)
        observation_ind(o, e, tick, i, out->obs_rel + slot * K, 0, NULL, NULL);
    if (out->coordinates) memcpy(out->coordinates + slot * 12, cbuf, sizeof(cbuf));
    if (out->agents) /* observation_handler.py:98 */
        for (int k = 0; k < N; ++k) {
            int ix, iy;
            coord2ind(o, en->xy[k], &ix, &iy);
            out->agents[(slot * N + k) * 2 + 0] = (int8_t)ix;
            out->agents[(slot * N + k) * 2 + 1] = (int8_t)iy;
        }
    if (out->objects) memcpy(out->objects + slot * T * XY, en->objects_matrix, T * XY); /* :99-101 */
    if (out->destination) memcpy(out->destination + slot * XY, en->destination + (size_t)i * XY, XY); /* :102-104 */
}

/* ------------------------------------------------------------------ reset */

/* default_environment.py:64-116 */
static void reset_one(orc_env *o, int e, uint32_t tick, const orc_out *out) {
    env_t *en = &o->envs[e];
    const int N = o->cfg.num_agents, T = o->cfg.type_objects;
    const size_t XY = (size_t)o->X * o->Y;
    memset(en->counters, 0, sizeof(en->counters));
    en->current_step = 0;
    en->pending_reset = 0;
    /* abstract_map.py:32-36 */
    memset(en->agents_matrix, 0, N * XY);
    memset(en->objects_matrix, 0, T * XY);
    reset_destination_area(o, e, tick);
    if (o->cfg.shuffle_init_xys) { /* :97-98 */
        uint32_t d = 0;
        const int n_init = o->cfg.num_init_xys > N ? o->cfg.num_init_xys : N; /* :33: one entry per init_xys row */
        tape_shuffle(o, e, tick, MAEXP_STREAM_INIT_PERM, &d, en->init_xys_order, n_init);
    }
    for (int i = 0; i < N; ++i) { /* :100-109 */
        const int ord = en->init_xys_order[i];
        en->collide_agents[i] = en->collide_walls[i] = 0;
        en->xy[i][0] = (int8_t)o->cfg.init_xys[ord][0];
        en->xy[i][1] = (int8_t)o->cfg.init_xys[ord][1];
        int ix, iy;
        coord2ind(o, en->xy[i], &ix, &iy);
        AM(o, en, i, ix, iy) = 1;
    }
    uint32_t attempt = 0; /* :112, :123-132; sequential_environment.py:21-23 / junctional_environment.py:21-23: type 0 only */
    for (int t = 0; t < (o->cfg.env_kind == ORC_ENV_DEFAULT ? T : 1); ++t)
        for (int n = 0; n < o->cfg.num_objects; ++n)
            generate_one_object(o, e, tick, MAEXP_STREAM_OBJECTS, &attempt, t);
    /* observation_handler.py:136-144 */
    if (uses_transition(o))
        for (int p = 0; p < o->cfg.past_step; ++p) get_past_global_agents(o, en, p);
    for (int i = 0; i < N; ++i) {
        emit_observation(o, e, tick, i, out);
        if (out && out->rewards) out->rewards[(size_t)e * N + i] = 0.0f;
        if (out && out->dones) out->dones[(size_t)e * N + i] = 0;
    }
    if (out && out->was_reset) out->was_reset[e] = 1;
}

void orc_reset(orc_env *o, uint32_t tick, const uint8_t *env_mask, const orc_out *out) {
    o->hm = out ? out->heatmaps : NULL;
    for (int e = 0; e < o->cfg.num_envs; ++e)
        if (!env_mask || env_mask[e]) reset_one(o, e, tick, out);
}

/* ------------------------------------------------------------------- step */

/* default_environment.py:153-175 */
static void step_one(orc_env *o, int e, uint32_t tick, const int8_t *actions, const uint8_t *order_in,
                     const orc_out *out) {
    env_t *en = &o->envs[e];
    const int N = o->cfg.num_agents, T = o->cfg.type_objects;
    if (o->cfg.auto_reset && en->pending_reset) { reset_one(o, e, tick, out); return; }

    /* :158-159, :180-194 */
    for (int i = 0; i < N; ++i) {
        int a;
        if (actions) a = actions[(size_t)e * N + i];
        else a = (int)(maexp_draw(o->cfg.seed, genv(o, e), tick, MAEXP_STREAM_ACTION, (uint32_t)i, 0) >> 30);
        static const int8_t delta[5][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}, {0, 0}};
        if (a >= 0 && a <= 4) { en->action[i][0] = delta[a][0]; en->action[i][1] = delta[a][1]; }
    }
    int32_t order[ORC_MAX_AGENTS];
    if (order_in) for (int j = 0; j < N; ++j) order[j] = order_in[(size_t)e * N + j];
    else {
        for (int j = 0; j < N; ++j) order[j] = j;
        uint32_t d = 0;
        tape_shuffle(o, e, tick, MAEXP_STREAM_ORDER, &d, order, N);
    }

    /* abstract_world.py:39-47, :54-72 */
    map_step(o, en); /* :40 */
    for (int j = 0; j < N; ++j) {
        const int id = order[j];
        int8_t force[2] = {en->action[id][0], en->action[id][1]};
        const int8_t next[2] = {(int8_t)(en->xy[id][0] + force[0]), (int8_t)(en->xy[id][1] + force[1])};
        for (int k = 0; k < N; ++k) {
            if (k == id) continue;
            if (next[0] == en->xy[k][0] && next[1] == en->xy[k][1]) {
                force[0] = force[1] = 0;
                en->collide_agents[id] = 1;
                break;
            }
        }
        int nx, ny;
        coord2ind(o, next, &nx, &ny);
        if (env_wall(o, en, nx, ny) == 1) { force[0] = force[1] = 0; en->collide_walls[id] = 1; }
        en->xy[id][0] = (int8_t)(en->xy[id][0] + force[0]);
        en->xy[id][1] = (int8_t)(en->xy[id][1] + force[1]);
    }
    /* :74-78 */
    memset(en->agents_matrix, 0, (size_t)N * o->X * o->Y);
    for (int i = 0; i < N; ++i) {
        int ix, iy;
        coord2ind(o, en->xy[i], &ix, &iy);
        AM(o, en, i, ix, iy) = 1;
    }

    /* default_environment.py:165-168 — reward, done, observation interleaved in agent order */
    float reward_n[ORC_MAX_AGENTS];
    uint8_t done_n[ORC_MAX_AGENTS];
    uint32_t attempt = 0;
    for (int i = 0; i < N; ++i) {
        int ax, ay;
        coord2ind(o, en->xy[i], &ax, &ay);
        float reward = 0.0f; /* :196-231 */
        if (o->hm) HM_AGENTS(o, i, ax, ay) += 1; /* :198 */
        if (o->cfg.task_mask[i] >= 0) { /* :201-202: task "-1" returns before the collision handling */
            for (int q = 0; q < ORC_MAX_TYPES && o->cfg.task_order[i][q] >= 0; ++q) {
                const int t = o->cfg.task_order[i][q];
                if (OM(o, en, t, ax, ay) == 1 && DM(o, en, i, ax, ay) == 1) {
                    reward = 1.0f;
                    OM(o, en, t, ax, ay) = 0;
                    en->counters[1] += 1;
                    if (o->hm) HM_COMPLETE(o, i, ax, ay) += 1; /* :213 */
                    if (o->cfg.env_kind == ORC_ENV_DEFAULT) {
                        if (o->cfg.keep_objects_num) generate_one_object(o, e, tick, MAEXP_STREAM_RESPAWN, &attempt, t);
                    } else {
                        /* sequential_environment.py:44-48, junctional_environment.py:44-56: the next object of the
                         * chain appears on the same cell (default_environment.py:118-121), the last one respawns type 0 */
                        int next = -1, respawn = 0;
                        if (o->cfg.env_kind == ORC_ENV_JUNCTIONAL && t == 1) {
                            if (i == 2 || i == 3) next = 3;
                            if (i == 4 || i == 5) next = 2;
                        } else if (o->cfg.env_kind == ORC_ENV_JUNCTIONAL && t == 2) next = 4;
                        else if (t < T - 1) next = t + 1;
                        else if (t == T - 1) respawn = o->cfg.keep_objects_num;
                        if (next >= 0) {
                            if (next >= T) abort(); /* IndexError in the reference */
                            OM(o, en, next, ax, ay) = 1;
                            if (o->hm) HM_OBJECTS(o, next, ax, ay) += 1; /* default_environment.py:118-121 */
                            en->counters[0] += 1;
                        }
                        if (respawn) generate_one_object(o, e, tick, MAEXP_STREAM_RESPAWN, &attempt, 0);
                    }
                }
            }
            if (en->collide_agents[i]) {
                reward = -1.0f; en->collide_agents[i] = 0; en->counters[2] += 1;
                if (o->hm) HM_COLLIDE(o, ax, ay) += 1; /* :221 */
            }
            if (en->collide_walls[i]) {
                reward = -1.0f; en->collide_walls[i] = 0; en->counters[3] += 1;
                if (o->hm) HM_WALL(o, ax, ay) += 1; /* :228 */
            }
        }
        reward_n[i] = reward;
        /* :233-237 */
        int total = 0;
        for (size_t z = 0; z < (size_t)T * o->X * o->Y; ++z) total += en->objects_matrix[z];
        done_n[i] = (total == 0 || en->current_step == o->cfg.max_episode_length - 1) ? 1 : 0;
        /* :239-248 */
        emit_observation(o, e, tick, i, out);
    }
    /* :170 → observation_handler.py:146-148 → transition_agent_observation.py:71-74 */
    if (uses_transition(o)) {
        const int P = o->cfg.past_step;
        const size_t frame = (size_t)N * o->X * o->Y;
        memmove(en->past_global_agents + frame, en->past_global_agents, frame * (size_t)(P - 1));
        get_past_global_agents(o, en, 0);
    }
    en->current_step += 1; /* :172 */
    if (o->hm) /* :173 */
        for (int t = 0; t < T; ++t)
            for (int x = 0; x < o->X; ++x)
                for (int y = 0; y < o->Y; ++y) HM_LEFT(o, t, x, y) += OM(o, en, t, x, y);

    /* shared_reward_environment.py:15-28 */
    if (o->cfg.reward_mode != ORC_REWARD_DEFAULT) {
        double s = 0.0, mx = reward_n[0];
        for (int i = 0; i < N; ++i) { s += (double)reward_n[i]; if (reward_n[i] > mx) mx = reward_n[i]; }
        const int mean = o->cfg.reward_mode == ORC_REWARD_SHARED_MEAN || o->cfg.shared_max_is_mean;
        const float shared = (float)(mean ? s / (double)N : mx);
        for (int i = 0; i < N; ++i) reward_n[i] = shared;
    }
    if (out && out->rewards) for (int i = 0; i < N; ++i) out->rewards[(size_t)e * N + i] = reward_n[i];
    if (out && out->dones) for (int i = 0; i < N; ++i) out->dones[(size_t)e * N + i] = done_n[i];
    if (out && out->was_reset) out->was_reset[e] = 0;
    if (o->cfg.auto_reset && done_n[0]) en->pending_reset = 1; /* default_trainer.py:85 returns dones[0] */
}

typedef struct {
    orc_env *o; uint32_t tick; const int8_t *actions; const uint8_t *order; const orc_out *out; int e0, e1;
} step_job;

static void *step_worker(void *p) {
    step_job *j = (step_job *)p;
    for (int e = j->e0; e < j->e1; ++e) step_one(j->o, e, j->tick, j->actions, j->order, j->out);
    return NULL;
}

/* envs are independent (no cross-env state in default_environment.py): contiguous slices per thread */
void orc_step(orc_env *o, uint32_t tick, const int8_t *actions, const uint8_t *order, const orc_out *out,
              int num_threads) {
    const int E = o->cfg.num_envs;
    o->hm = out ? out->heatmaps : NULL;
    if (o->hm) num_threads = 1; /* the shared statistics are plain (non-atomic) sums */
    if (num_threads > 64) num_threads = 64;
    if (num_threads > E) num_threads = E;
    if (num_threads <= 1) {
        for (int e = 0; e < E; ++e) step_one(o, e, tick, actions, order, out);
        return;
    }
    pthread_t th[64];
    step_job jobs[64];
    for (int t = 0; t < num_threads; ++t) {
        jobs[t] = (step_job){o, tick, actions, order, out, (int)((long long)E * t / num_threads),
                             (int)((long long)E * (t + 1) / num_threads)};
        pthread_create(&th[t], NULL, step_worker, &jobs[t]);
    }
    for (int t = 0; t < num_threads; ++t) pthread_join(th[t], NULL);
}

void orc_get_state(const orc_env *o, int8_t *pos, int8_t *objects, int32_t *counters, int32_t *steps,
                   int8_t *destination) {
    const int N = o->cfg.num_agents, T = o->cfg.type_objects;
    const size_t XY = (size_t)o->X * o->Y;
    for (int e = 0; e < o->cfg.num_envs; ++e) {
        const env_t *en = &o->envs[e];
        if (pos)
            for (int i = 0; i < N; ++i) {
                int ix, iy;
                coord2ind(o, en->xy[i], &ix, &iy);
                pos[((size_t)e * N + i) * 2 + 0] = (int8_t)ix;
                pos[((size_t)e * N + i) * 2 + 1] = (int8_t)iy;
            }
        if (objects) memcpy(objects + (size_t)e * T * XY, en->objects_matrix, T * XY);
        if (counters) memcpy(counters + (size_t)e * 4, en->counters, sizeof(en->counters));
        if (steps) steps[e] = en->current_step;
        if (destination) memcpy(destination + (size_t)e * N * XY, en->destination, N * XY);
    }
}

void orc_set_state(orc_env *o, const int8_t *pos, const int8_t *objects) {
    const int N = o->cfg.num_agents, T = o->cfg.type_objects;
    const size_t XY = (size_t)o->X * o->Y;
    for (int e = 0; e < o->cfg.num_envs; ++e) {
        env_t *en = &o->envs[e];
        if (pos) {
            memset(en->agents_matrix, 0, N * XY);
            for (int i = 0; i < N; ++i) {
                const int ix = pos[((size_t)e * N + i) * 2 + 0], iy = pos[((size_t)e * N + i) * 2 + 1];
                en->xy[i][0] = (int8_t)(ix - o->X / 2);
                en->xy[i][1] = (int8_t)(o->Y / 2 - iy);
                AM(o, en, i, ix, iy) = 1;
            }
        }
        if (objects) memcpy(en->objects_matrix, objects + (size_t)e * T * XY, T * XY);
    }
}
