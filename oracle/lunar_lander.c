/* oracle/lunar_lander.c — CPU restatement of LunarLander-v3 (TEST INFRASTRUCTURE, not product code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product path (gymrl_b200/csrc/env_lunar.cu) never does.
 *
 * PARITY UNPINNED: the reference (Starlight0798/gymRL) only *calls* gym.make("LunarLander-v3")
 * (algorithms/ppo_lunarlander.py:160,200,211,222; ppo_full_lunarlander.py:443,466,478,496); the
 * arithmetic lives in third-party `gymnasium` (un-pinned, requirements.txt:5-6; v3 implies >= 1.0)
 * and `box2d-py` (Box2D 2.3.x), neither of which is in /root/reference nor installable here, and
 * the reference has no tests or golden vectors for it.  This file restates
 *   - gymnasium/envs/box2d/lunar_lander.py (LunarLander.reset/step, ContactDetector), and
 *   - Box2D 2.3's published algorithm as used by that file: b2PolygonShape::Set/ComputeMass,
 *     b2CollideEdgeAndPolygon (isolated edges), b2ContactSolver (warm start, block solver,
 *     Baumgarte position correction), b2RevoluteJoint (motor + limits), b2Island::Solve
 *     (180 velocity / 60 position iterations, sleeping).
 * Documented deviations: no TOI sub-stepping (b2World::SolveTOI), a fixed
 * contact order (body-major, edge-minor) instead of Box2D's list order, at most LL_MAX_MANIFOLDS
 * touching manifolds per env, Philox instead of PCG64 for np_random, and a fixed-sequence
 * polynomial sin/cos (so that the CUDA kernel and this file agree bit-for-bit).
 *
 * All solver arithmetic is float32 in the same operation order as the device kernel; compile with
 * -ffp-contract=off (oracle/Makefile does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FPS 50
#define SCALE 30.0
#define MAIN_ENGINE_POWER 13.0
#define SIDE_ENGINE_POWER 0.6
#define INITIAL_RANDOM 1000.0
#define LEG_AWAY 20
#define LEG_DOWN 18
#define LEG_W 2
#define LEG_H 8
#define LEG_SPRING_TORQUE 40
#define SIDE_ENGINE_HEIGHT 14
#define SIDE_ENGINE_AWAY 12
#define MAIN_ENGINE_Y_LOCATION 4
#define VIEWPORT_W 600
#define VIEWPORT_H 400
#define CHUNKS 11
#define MAX_EPISODE_STEPS 1000

#define B2_LINEAR_SLOP 0.005f
#define B2_ANGULAR_SLOP (2.0f / 180.0f * 3.14159265359f)
#define B2_POLYGON_RADIUS (2.0f * B2_LINEAR_SLOP)
#define B2_MAX_LINEAR_CORRECTION 0.2f
#define B2_MAX_ANGULAR_CORRECTION (8.0f / 180.0f * 3.14159265359f)
#define B2_BAUMGARTE 0.2f
#define B2_VELOCITY_THRESHOLD 1.0f
#define B2_MAX_TRANSLATION 2.0f
#define B2_MAX_ROTATION (0.5f * 3.14159265359f)
#define B2_TIME_TO_SLEEP 0.5f
#define B2_LINEAR_SLEEP_TOL 0.01f
#define B2_ANGULAR_SLEEP_TOL (2.0f / 180.0f * 3.14159265359f)
#define VEL_ITERS 180
#define POS_ITERS 60

#define LL_NBODY 3  /* 0 lander, 1 legs[0] (i=-1), 2 legs[1] (i=+1) */
#define LL_NEDGE 11 /* 10 terrain edges + the (0,0)-(W,0) base edge */
#define LL_MAX_MANIFOLDS 8
#define LL_STATE_DOUBLES 128

typedef struct { float x, y; } v2;

static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 add(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 sub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 neg(v2 a) { return V(-a.x, -a.y); }
static inline v2 mul(float s, v2 a) { return V(s * a.x, s * a.y); }
static inline float dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float cross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline v2 cross_vs(v2 a, float s) { return V(s * a.y, -s * a.x); }  /* b2Cross(v, s) */
static inline v2 cross_sv(float s, v2 a) { return V(-s * a.y, s * a.x); }  /* b2Cross(s, v) */
static inline float clampf(float a, float lo, float hi) { return fmaxf(lo, fminf(a, hi)); }

/* Fused helpers of the velocity iterations (the step's serial chain): explicit fmaf, so that the CUDA kernel (built with
 * -fmad=false) and this file (built with -ffp-contract=off) fuse exactly the same multiply-adds and round alike. */
static inline float fdot(v2 a, v2 b) { return fmaf(a.x, b.x, a.y * b.y); }
static inline float fcross(v2 a, v2 b) { return fmaf(a.x, b.y, -(a.y * b.x)); }
static inline v2 axpy(float s, v2 x, v2 y) { return V(fmaf(s, x.x, y.x), fmaf(s, x.y, y.y)); }                 /* y + s x */
static inline v2 add_cross_sv(v2 a, float s, v2 r) { return V(fmaf(-s, r.y, a.x), fmaf(s, r.x, a.y)); }       /* a + b2Cross(s, r) */
static inline v2 sub_cross_sv(v2 a, float s, v2 r) { return V(fmaf(s, r.y, a.x), fmaf(-s, r.x, a.y)); }       /* a - b2Cross(s, r) */

/* ---- fixed-sequence sin/cos (Cody-Waite pi/2 reduction + cephes minimax polynomials) ---------- */
static void det_sincosf(float a, float* s, float* c) {
    const float k = rintf(a * 0.636619772367581343f);
    float r = a - k * 1.5703125f;
    r = r - k * 4.837512969970703125e-4f;
    r = r - k * 7.54978995489188e-8f;
    const float z = r * r;
    float sp = -1.9515295891e-4f * z + 8.3321608736e-3f;
    sp = sp * z - 1.6666654611e-1f;
    sp = sp * z * r + r;
    float cp = 2.443315711809948e-5f * z - 1.388731625493765e-3f;
    cp = cp * z + 4.166664568298827e-2f;
    cp = cp * z * z - 0.5f * z + 1.0f;
    const int q = ((int)k) & 3;
    if (q == 0) { *s = sp; *c = cp; }
    else if (q == 1) { *s = cp; *c = -sp; }
    else if (q == 2) { *s = -sp; *c = -cp; }
    else { *s = -cp; *c = sp; }
}

typedef struct { float s, c; } rot;
static inline rot make_rot(float a) { rot q; det_sincosf(a, &q.s, &q.c); return q; }
static inline v2 rmul(rot q, v2 v) { return V(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
static inline v2 rmulT(rot q, v2 v) { return V(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }

/* ---- Philox4x32-10 (independent restatement; must match gymrl_b200/csrc/common.cuh) ------------ */
static void philox(uint64_t seed, uint64_t entity, uint32_t draw, uint32_t stream, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)entity, c1 = draw, c2 = stream, c3 = (uint32_t)(entity >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u01(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

/* ---- shapes ------------------------------------------------------------------------------------ */
typedef struct {
    int count;
    v2 v[6], n[6];
    v2 centroid;
} Poly;

typedef struct {
    Poly poly[LL_NBODY];
    float inv_mass[LL_NBODY], inv_I[LL_NBODY];
    v2 local_center[LL_NBODY];
    float friction[LL_NBODY]; /* mixed with the terrain's 0.1 (or the base edge's 0.2) at solve time */
    int ready;
} Shapes;
static Shapes g_shapes;

static void poly_finish(Poly* p) {
    for (int i = 0; i < p->count; ++i) {
        int j = i + 1 < p->count ? i + 1 : 0;
        v2 e = sub(p->v[j], p->v[i]);
        v2 nn = cross_vs(e, 1.0f);
        float len = sqrtf(nn.x * nn.x + nn.y * nn.y);
        float inv = 1.0f / len;
        p->n[i] = V(nn.x * inv, nn.y * inv);
    }
}

/* b2PolygonShape::ComputeCentroid (reference point at the origin) */
static v2 poly_centroid(const Poly* p) {
    v2 c = V(0.0f, 0.0f);
    float area = 0.0f;
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < p->count; ++i) {
        v2 p2 = p->v[i], p3 = p->v[i + 1 < p->count ? i + 1 : 0];
        float D = cross(p2, p3);
        float ta = 0.5f * D;
        area += ta;
        c = add(c, mul(ta * inv3, add(p2, p3)));
    }
    return mul(1.0f / area, c);
}

/* b2PolygonShape::ComputeMass followed by b2Body::ResetMassData for a single-fixture body */
static void poly_mass(const Poly* p, float density, float* inv_mass, float* inv_I, v2* local_center) {
    v2 center = V(0.0f, 0.0f), s = V(0.0f, 0.0f);
    float area = 0.0f, I = 0.0f;
    for (int i = 0; i < p->count; ++i) s = add(s, p->v[i]);
    s = mul(1.0f / (float)p->count, s);
    const float k_inv3 = 1.0f / 3.0f;
    for (int i = 0; i < p->count; ++i) {
        v2 e1 = sub(p->v[i], s), e2 = sub(p->v[i + 1 < p->count ? i + 1 : 0], s);
        float D = cross(e1, e2);
        float ta = 0.5f * D;
        area += ta;
        center = add(center, mul(ta * k_inv3, add(e1, e2)));
        float intx2 = e1.x * e1.x + e2.x * e1.x + e2.x * e2.x;
        float inty2 = e1.y * e1.y + e2.y * e1.y + e2.y * e2.y;
        I += (0.25f * k_inv3 * D) * (intx2 + inty2);
    }
    float mass = density * area;
    center = mul(1.0f / area, center);
    v2 mc = add(center, s);
    float Io = density * I;
    Io += mass * (dot(mc, mc) - dot(center, center));
    /* ResetMassData */
    float m = mass;
    v2 lc = mul(1.0f / m, mul(mass, mc));
    float Ic = Io - m * dot(lc, lc);
    *inv_mass = 1.0f / m;
    *inv_I = 1.0f / Ic;
    *local_center = lc;
}

static void shapes_init(void) {
    if (g_shapes.ready) return;
    /* LANDER_POLY after b2PolygonShape::Set's gift wrapping: starts at the right-most lowest vertex, CCW. */
    const double lp[6][2] = {{17, -10}, {17, 0}, {14, 17}, {-14, 17}, {-17, 0}, {-17, -10}};
    Poly* L = &g_shapes.poly[0];
    L->count = 6;
    for (int i = 0; i < 6; ++i) L->v[i] = V((float)(lp[i][0] / SCALE), (float)(lp[i][1] / SCALE));
    poly_finish(L);
    L->centroid = poly_centroid(L);
    poly_mass(L, 5.0f, &g_shapes.inv_mass[0], &g_shapes.inv_I[0], &g_shapes.local_center[0]);
    g_shapes.friction[0] = 0.1f;
    for (int b = 1; b < LL_NBODY; ++b) { /* SetAsBox(LEG_W/SCALE, LEG_H/SCALE) */
        Poly* P = &g_shapes.poly[b];
        float hx = (float)(LEG_W / SCALE), hy = (float)(LEG_H / SCALE);
        P->count = 4;
        P->v[0] = V(-hx, -hy); P->v[1] = V(hx, -hy); P->v[2] = V(hx, hy); P->v[3] = V(-hx, hy);
        P->n[0] = V(0.0f, -1.0f); P->n[1] = V(1.0f, 0.0f); P->n[2] = V(0.0f, 1.0f); P->n[3] = V(-1.0f, 0.0f);
        P->centroid = V(0.0f, 0.0f);
        poly_mass(P, 1.0f, &g_shapes.inv_mass[b], &g_shapes.inv_I[b], &g_shapes.local_center[b]);
        g_shapes.friction[b] = 0.2f; /* b2FixtureDef default */
    }
    g_shapes.ready = 1;
}

/* ---- world state ------------------------------------------------------------------------------- */
typedef struct {
    v2 c;          /* centre of mass (b2Sweep::c) */
    float a;       /* angle */
    v2 v;
    float w;
    float sleep_time;
} Body;

typedef struct {
    float imp_x, imp_y, imp_z, motor_impulse;
    int limit_state; /* 0 inactive, 1 at lower, 2 at upper, 3 equal */
} Joint;

typedef struct {
    int key;   /* body*16 + edge, -1 = free */
    int count;
    uint32_t id[2];
    float nimp[2], timp[2];
} ManifoldSlot;

typedef struct LLEnv {
    uint64_t seed, env_id;
    float terrain_y[CHUNKS]; /* smooth_y as float32 edge vertices */
    Body body[LL_NBODY];
    Joint joint[2];
    ManifoldSlot slot[LL_MAX_MANIFOLDS];
    v2 pending_force; /* ApplyForceToCenter at reset, consumed by the first Step */
    int game_over, leg_contact[2], awake;
    int has_prev_shaping;
    double prev_shaping;
    int elapsed;
    uint32_t episode, stepctr;
    double ep_return;
} LLEnv;

static const double W_ = VIEWPORT_W / SCALE, H_ = VIEWPORT_H / SCALE;

static float chunk_x(int i) { return (float)(W_ / (CHUNKS - 1) * i); }

static void edge_verts(const LLEnv* e, int k, v2* v1, v2* v2_) {
    if (k < CHUNKS - 1) {
        *v1 = V(chunk_x(k), e->terrain_y[k]);
        *v2_ = V(chunk_x(k + 1), e->terrain_y[k + 1]);
    } else {
        *v1 = V(0.0f, 0.0f);
        *v2_ = V((float)W_, 0.0f);
    }
}

/* joint definitions: bodyA = lander, bodyB = leg; anchors in body-local frames */
static const float JOINT_SIGN[2] = {-1.0f, +1.0f};
static v2 joint_anchor_b(int j) { return V((float)(JOINT_SIGN[j] * LEG_AWAY / SCALE), (float)(LEG_DOWN / SCALE)); }
static float joint_lower(int j) { return j == 0 ? (float)(+0.9 - 0.5) : (float)(-0.9); }
static float joint_upper(int j) { return j == 0 ? (float)(+0.9) : (float)(-0.9 + 0.5); }
static float joint_motor_speed(int j) { return (float)(+0.3 * JOINT_SIGN[j]); }
/* pybox2d's b2RevoluteJointDef(**kw) sets referenceAngle = bodyB.angle - bodyA.angle when it is not
 * passed explicitly; the legs are created at angle i*0.05 and the lander at 0. */
static float joint_ref_angle(int j) { return (float)(JOINT_SIGN[j] * 0.05) - 0.0f; }

/* ---- collision: b2CollideEdgeAndPolygon for an edge without adjacent vertices ------------------ */
typedef struct {
    int count;
    int type; /* 0 faceA (edge is reference), 1 faceB (polygon face is reference) */
    v2 local_normal, local_point;
    v2 pt[2]; /* faceA: in polygon-local frame; faceB: in edge (world) frame */
    uint32_t id[2];
} Manifold;

typedef struct { v2 v; uint8_t ia, ib, ta, tb; } ClipVertex;
static inline uint32_t cf_key(uint8_t ia, uint8_t ib, uint8_t ta, uint8_t tb) {
    return (uint32_t)ia | ((uint32_t)ib << 8) | ((uint32_t)ta << 16) | ((uint32_t)tb << 24);
}

static int clip_segment(ClipVertex out[2], const ClipVertex in[2], v2 normal, float offset, int vertexIndexA) {
    int n = 0;
    float d0 = dot(normal, in[0].v) - offset;
    float d1 = dot(normal, in[1].v) - offset;
    if (d0 <= 0.0f) out[n++] = in[0];
    if (d1 <= 0.0f) out[n++] = in[1];
    if (d0 * d1 < 0.0f) {
        float interp = d0 / (d0 - d1);
        out[n].v = add(in[0].v, mul(interp, sub(in[1].v, in[0].v)));
        out[n].ia = (uint8_t)vertexIndexA;
        out[n].ib = in[0].ib;
        out[n].ta = 0; /* e_vertex */
        out[n].tb = 1; /* e_face */
        ++n;
    }
    return n;
}

/* xfB = (p, q) of the polygon body; the edge lives on the static body at the identity transform. */
static void collide_edge_polygon(Manifold* m, v2 ev1, v2 ev2, const Poly* poly, v2 xp, rot xq) {
    m->count = 0;
    v2 centroidB = add(rmul(xq, poly->centroid), xp);
    v2 edge1 = sub(ev2, ev1);
    {
        float len = sqrtf(edge1.x * edge1.x + edge1.y * edge1.y);
        float inv = 1.0f / len;
        edge1 = V(edge1.x * inv, edge1.y * inv);
    }
    v2 normal1 = V(edge1.y, -edge1.x);
    float offset1 = dot(normal1, sub(centroidB, ev1));
    int front = offset1 >= 0.0f;
    v2 normal = front ? normal1 : neg(normal1);

    v2 pv[6], pn[6];
    for (int i = 0; i < poly->count; ++i) {
        pv[i] = add(rmul(xq, poly->v[i]), xp);
        pn[i] = rmul(xq, poly->n[i]);
    }
    const float radius = 2.0f * B2_POLYGON_RADIUS;

    /* ComputeEdgeSeparation */
    float edge_sep = 3.402823466e+38f;
    for (int i = 0; i < poly->count; ++i) {
        float s = dot(normal, sub(pv[i], ev1));
        if (s < edge_sep) edge_sep = s;
    }
    if (edge_sep > radius) return;

    /* ComputePolygonSeparation (upper/lower limit = -normal: the adjacency test never rejects) */
    int poly_type = 0, poly_index = -1; /* 0 unknown, 2 edgeB */
    float poly_sep = -3.402823466e+38f;
    for (int i = 0; i < poly->count; ++i) {
        v2 n = neg(pn[i]);
        float s1 = dot(n, sub(pv[i], ev1));
        float s2 = dot(n, sub(pv[i], ev2));
        float s = fminf(s1, s2);
        if (s > radius) { poly_type = 2; poly_index = i; poly_sep = s; break; }
        if (s > poly_sep) { poly_type = 2; poly_index = i; poly_sep = s; }
    }
    if (poly_type != 0 && poly_sep > radius) return;

    int primary_is_edge;
    if (poly_type == 0) primary_is_edge = 1;
    else if (poly_sep > 0.98f * edge_sep + 0.001f) primary_is_edge = 0;
    else primary_is_edge = 1;

    ClipVertex ie[2];
    int rf_i1, rf_i2;
    v2 rf_v1, rf_v2, rf_normal;
    if (primary_is_edge) {
        m->type = 0;
        int best = 0;
        float best_val = dot(normal, pn[0]);
        for (int i = 1; i < poly->count; ++i) {
            float val = dot(normal, pn[i]);
            if (val < best_val) { best_val = val; best = i; }
        }
        int i1 = best, i2 = i1 + 1 < poly->count ? i1 + 1 : 0;
        ie[0].v = pv[i1]; ie[0].ia = 0; ie[0].ib = (uint8_t)i1; ie[0].ta = 1; ie[0].tb = 0;
        ie[1].v = pv[i2]; ie[1].ia = 0; ie[1].ib = (uint8_t)i2; ie[1].ta = 1; ie[1].tb = 0;
        if (front) { rf_i1 = 0; rf_i2 = 1; rf_v1 = ev1; rf_v2 = ev2; rf_normal = normal1; }
        else { rf_i1 = 1; rf_i2 = 0; rf_v1 = ev2; rf_v2 = ev1; rf_normal = neg(normal1); }
    } else {
        m->type = 1;
        ie[0].v = ev1; ie[0].ia = 0; ie[0].ib = (uint8_t)poly_index; ie[0].ta = 0; ie[0].tb = 1;
        ie[1].v = ev2; ie[1].ia = 0; ie[1].ib = (uint8_t)poly_index; ie[1].ta = 0; ie[1].tb = 1;
        rf_i1 = poly_index;
        rf_i2 = rf_i1 + 1 < poly->count ? rf_i1 + 1 : 0;
        rf_v1 = pv[rf_i1]; rf_v2 = pv[rf_i2]; rf_normal = pn[rf_i1];
    }
    v2 side1 = V(rf_normal.y, -rf_normal.x), side2 = neg(side1);
    float off1 = dot(side1, rf_v1), off2 = dot(side2, rf_v2);
    ClipVertex cp1[2], cp2[2];
    if (clip_segment(cp1, ie, side1, off1, rf_i1) < 2) return;
    if (clip_segment(cp2, cp1, side2, off2, rf_i2) < 2) return;

    if (primary_is_edge) { m->local_normal = rf_normal; m->local_point = rf_v1; }
    else { m->local_normal = poly->n[rf_i1]; m->local_point = poly->v[rf_i1]; }

    int pc = 0;
    for (int i = 0; i < 2; ++i) {
        float sep = dot(rf_normal, sub(cp2[i].v, rf_v1));
        if (sep <= radius) {
            if (primary_is_edge) {
                m->pt[pc] = rmulT(xq, sub(cp2[i].v, xp));
                m->id[pc] = cf_key(cp2[i].ia, cp2[i].ib, cp2[i].ta, cp2[i].tb);
            } else {
                m->pt[pc] = cp2[i].v;
                m->id[pc] = cf_key(cp2[i].ib, cp2[i].ia, cp2[i].tb, cp2[i].ta);
            }
            ++pc;
        }
    }
    m->count = pc;
}

/* ---- contact constraints ------------------------------------------------------------------------ */
typedef struct {
    int body, edge;
    Manifold man;
    float friction;
    /* velocity constraint */
    v2 normal;
    int vc_count;
    v2 rB[2];
    float normal_mass[2], tangent_mass[2], velocity_bias[2];
    float nimp[2], timp[2];
    float K11, K12, K22, NM11, NM12, NM21, NM22; /* K and its inverse (block solver) */
} Contact;

static void body_xf(const Body* b, int bi, v2* p, rot* q) {
    *q = make_rot(b->a);
    *p = sub(b->c, rmul(*q, g_shapes.local_center[bi]));
}

/* b2WorldManifold::Initialize with A = edge (identity xf, radius r), B = polygon */
static void world_manifold(const Manifold* m, v2 xpB, rot xqB, v2* normal, v2 pts[2]) {
    const float rA = B2_POLYGON_RADIUS, rB = B2_POLYGON_RADIUS;
    if (m->type == 0) {
        *normal = m->local_normal;
        v2 plane = m->local_point;
        for (int i = 0; i < m->count; ++i) {
            v2 clip = add(rmul(xqB, m->pt[i]), xpB);
            v2 cA = add(clip, mul(rA - dot(sub(clip, plane), *normal), *normal));
            v2 cB = sub(clip, mul(rB, *normal));
            pts[i] = mul(0.5f, add(cA, cB));
        }
    } else {
        v2 n = rmul(xqB, m->local_normal);
        v2 plane = add(rmul(xqB, m->local_point), xpB);
        for (int i = 0; i < m->count; ++i) {
            v2 clip = m->pt[i];
            v2 cB = add(clip, mul(rB - dot(sub(clip, plane), n), n));
            v2 cA = sub(clip, mul(rA, n));
            pts[i] = mul(0.5f, add(cA, cB));
        }
        *normal = neg(n);
    }
}

/* ---- the environment ---------------------------------------------------------------------------- */
static void ll_world_step(LLEnv* e);
static void ll_observe(LLEnv* e, double st[8]);

static void ll_begin_episode(LLEnv* e) {
    shapes_init();
    uint32_t r[28];
    for (int j = 0; j < 7; ++j) philox(e->seed, e->env_id, e->episode * 8u + (uint32_t)j, 0u, r + 4 * j);
    double height[CHUNKS + 1];
    for (int i = 0; i <= CHUNKS; ++i) height[i] = 0.0 + (H_ / 2 - 0.0) * u01(r[2 * i], r[2 * i + 1]);
    double fx = -INITIAL_RANDOM + (INITIAL_RANDOM - -INITIAL_RANDOM) * u01(r[24], r[25]);
    double fy = -INITIAL_RANDOM + (INITIAL_RANDOM - -INITIAL_RANDOM) * u01(r[26], r[27]);
    e->episode += 1;

    const double helipad_y = H_ / 4;
    for (int k = -2; k <= 2; ++k) height[CHUNKS / 2 + k] = helipad_y;
    for (int i = 0; i < CHUNKS; ++i) {
        /* height[i-1] with i = 0 wraps to height[-1] == height[CHUNKS] (Python negative index) */
        double hm = height[i == 0 ? CHUNKS : i - 1];
        e->terrain_y[i] = (float)(0.33 * (hm + height[i + 0] + height[i + 1]));
    }
    const float ix = (float)(VIEWPORT_W / SCALE / 2), iy = (float)(VIEWPORT_H / SCALE);
    for (int b = 0; b < LL_NBODY; ++b) {
        Body* B = &e->body[b];
        float ang = 0.0f;
        v2 pos = V(ix, iy);
        if (b > 0) {
            double i = JOINT_SIGN[b - 1];
            pos = V((float)((double)ix - i * LEG_AWAY / SCALE), iy);
            ang = (float)(i * 0.05);
        }
        rot q = make_rot(ang);
        B->a = ang;
        B->c = add(rmul(q, g_shapes.local_center[b]), pos);
        B->v = V(0.0f, 0.0f);
        B->w = 0.0f;
        B->sleep_time = 0.0f;
    }
    for (int j = 0; j < 2; ++j) {
        e->joint[j].imp_x = e->joint[j].imp_y = e->joint[j].imp_z = e->joint[j].motor_impulse = 0.0f;
        e->joint[j].limit_state = 0;
    }
    for (int s = 0; s < LL_MAX_MANIFOLDS; ++s) { memset(&e->slot[s], 0, sizeof(ManifoldSlot)); e->slot[s].key = -1; }
    e->pending_force = V((float)fx, (float)fy);
    e->game_over = 0;
    e->leg_contact[0] = e->leg_contact[1] = 0;
    e->awake = 1;
    e->has_prev_shaping = 0;
    e->prev_shaping = 0.0;
    e->elapsed = 0;
    e->ep_return = 0.0;
}

/* One LunarLander.step(action): engines -> world.Step -> observation / reward / termination.
 * Returns the reward; st[] receives the float64 `state` list of lunar_lander.py. */
static double ll_env_step(LLEnv* e, int action, double st[8], int* terminated) {
    Body* L = &e->body[0];
    float sa, ca;
    det_sincosf(L->a, &sa, &ca);
    const double tip0 = (double)sa, tip1 = (double)ca;
    const double side0 = -tip1, side1 = tip0;
    uint32_t r[4];
    philox(e->seed, e->env_id, e->stepctr, 1u, r);
    e->stepctr += 1;
    const double disp0 = (-1.0 + 2.0 * u01(r[0], r[1])) / SCALE;
    const double disp1 = (-1.0 + 2.0 * u01(r[2], r[3])) / SCALE;
    rot q = make_rot(L->a);
    v2 lpos = sub(L->c, rmul(q, g_shapes.local_center[0])); /* lander.position (body origin) */

    double m_power = 0.0, s_power = 0.0;
    if (action == 2) {
        m_power = 1.0;
        double ox = tip0 * (MAIN_ENGINE_Y_LOCATION / SCALE + 2 * disp0) + side0 * disp1;
        double oy = -tip1 * (MAIN_ENGINE_Y_LOCATION / SCALE + 2 * disp0) - side1 * disp1;
        v2 ip = V((float)((double)lpos.x + ox), (float)((double)lpos.y + oy));
        v2 imp = V((float)(-ox * MAIN_ENGINE_POWER * m_power), (float)(-oy * MAIN_ENGINE_POWER * m_power));
        L->v = add(L->v, mul(g_shapes.inv_mass[0], imp));
        L->w += g_shapes.inv_I[0] * cross(sub(ip, L->c), imp);
    }
    if (action == 1 || action == 3) {
        double direction = (double)(action - 2);
        s_power = 1.0;
        double ox = tip0 * disp0 + side0 * (3 * disp1 + direction * SIDE_ENGINE_AWAY / SCALE);
        double oy = -tip1 * disp0 - side1 * (3 * disp1 + direction * SIDE_ENGINE_AWAY / SCALE);
        v2 ip = V((float)((double)lpos.x + ox - tip0 * 17 / SCALE),
                  (float)((double)lpos.y + oy + tip1 * SIDE_ENGINE_HEIGHT / SCALE));
        v2 imp = V((float)(-ox * SIDE_ENGINE_POWER * s_power), (float)(-oy * SIDE_ENGINE_POWER * s_power));
        L->v = add(L->v, mul(g_shapes.inv_mass[0], imp));
        L->w += g_shapes.inv_I[0] * cross(sub(ip, L->c), imp);
    }

    ll_world_step(e);
    ll_observe(e, st);

    double reward = 0.0;
    double shaping = -100 * sqrt(st[0] * st[0] + st[1] * st[1]) - 100 * sqrt(st[2] * st[2] + st[3] * st[3]) -
                     100 * fabs(st[4]) + 10 * st[6] + 10 * st[7];
    if (e->has_prev_shaping) reward = shaping - e->prev_shaping;
    e->prev_shaping = shaping;
    e->has_prev_shaping = 1;
    reward -= m_power * 0.30;
    reward -= s_power * 0.03;
    *terminated = 0;
    if (e->game_over || fabs(st[0]) >= 1.0) { *terminated = 1; reward = -100; }
    if (!e->awake) { *terminated = 1; reward = +100; }
    return reward;
}

static void ll_observe(LLEnv* e, double st[8]) {
    const Body* L = &e->body[0];
    rot q = make_rot(L->a);
    v2 pos = sub(L->c, rmul(q, g_shapes.local_center[0]));
    const double helipad_y = H_ / 4;
    st[0] = ((double)pos.x - VIEWPORT_W / SCALE / 2) / (VIEWPORT_W / SCALE / 2);
    st[1] = ((double)pos.y - (helipad_y + LEG_DOWN / SCALE)) / (VIEWPORT_H / SCALE / 2);
    st[2] = (double)L->v.x * (VIEWPORT_W / SCALE / 2) / FPS;
    st[3] = (double)L->v.y * (VIEWPORT_H / SCALE / 2) / FPS;
    st[4] = (double)L->a;
    st[5] = 20.0 * (double)L->w / FPS;
    st[6] = e->leg_contact[0] ? 1.0 : 0.0;
    st[7] = e->leg_contact[1] ? 1.0 : 0.0;
}

/* ---- b2World::Step(1/50, 180, 60): Collide + one-island Solve ---------------------------------- */
static void ll_world_step(LLEnv* e) {
    const float h = (float)(1.0 / FPS);
    const float gx = 0.0f, gy = -10.0f;
    Contact con[LL_MAX_MANIFOLDS];
    int nc = 0;

    /* Collide: evaluate every (polygon, edge) pair; carry impulses over by contact-feature id;
     * fire Begin/EndContact on touching transitions (ContactDetector). */
    ManifoldSlot new_slot[LL_MAX_MANIFOLDS];
    for (int s = 0; s < LL_MAX_MANIFOLDS; ++s) { memset(&new_slot[s], 0, sizeof(ManifoldSlot)); new_slot[s].key = -1; }
    for (int b = 0; b < LL_NBODY; ++b) {
        v2 xp; rot xq;
        body_xf(&e->body[b], b, &xp, &xq);
        /* polygon AABB in world space (conservative pre-test; does not change results) */
        float minx = 3.4e38f, maxx = -3.4e38f, miny = 3.4e38f, maxy = -3.4e38f;
        for (int i = 0; i < g_shapes.poly[b].count; ++i) {
            v2 p = add(rmul(xq, g_shapes.poly[b].v[i]), xp);
            minx = fminf(minx, p.x); maxx = fmaxf(maxx, p.x);
            miny = fminf(miny, p.y); maxy = fmaxf(maxy, p.y);
        }
        for (int k = 0; k < LL_NEDGE; ++k) {
            const int key = b * 16 + k;
            int old = -1;
            for (int s = 0; s < LL_MAX_MANIFOLDS; ++s) if (e->slot[s].key == key) old = s;
            v2 ev1, ev2;
            edge_verts(e, k, &ev1, &ev2);
            Manifold m;
            m.count = 0;
            const float margin = 0.1f;
            if (!(maxx + margin < fminf(ev1.x, ev2.x) || minx - margin > fmaxf(ev1.x, ev2.x) ||
                  maxy + margin < fminf(ev1.y, ev2.y) || miny - margin > fmaxf(ev1.y, ev2.y)))
                collide_edge_polygon(&m, ev1, ev2, &g_shapes.poly[b], xp, xq);
            int touching = m.count > 0 && nc < LL_MAX_MANIFOLDS;
            int was = old >= 0;
            if (touching && !was) { /* BeginContact */
                if (b == 0) e->game_over = 1; else e->leg_contact[b - 1] = 1;
            }
            if (!touching && was) { /* EndContact */
                if (b > 0) e->leg_contact[b - 1] = 0;
            }
            if (touching) {
                Contact* c = &con[nc];
                c->body = b; c->edge = k; c->man = m;
                float fe = k < CHUNKS - 1 ? 0.1f : 0.2f;
                c->friction = sqrtf(g_shapes.friction[b] * fe);
                for (int i = 0; i < m.count; ++i) {
                    c->nimp[i] = 0.0f; c->timp[i] = 0.0f;
                    if (was)
                        for (int j = 0; j < e->slot[old].count; ++j)
                            if (e->slot[old].id[j] == m.id[i]) { c->nimp[i] = e->slot[old].nimp[j]; c->timp[i] = e->slot[old].timp[j]; break; }
                }
                ++nc;
            }
        }
    }

    /* ---- b2Island::Solve ---- */
    Body* B = e->body;
    const float* im = g_shapes.inv_mass;
    const float* ii = g_shapes.inv_I;
    /* integrate velocities (gravity, the one-shot reset force; damping is zero) */
    for (int b = 0; b < LL_NBODY; ++b) {
        v2 f = b == 0 ? e->pending_force : V(0.0f, 0.0f);
        B[b].v.x += h * (gx + im[b] * f.x);
        B[b].v.y += h * (gy + im[b] * f.y);
        B[b].v = mul(1.0f / (1.0f + h * 0.0f), B[b].v);
        B[b].w *= 1.0f / (1.0f + h * 0.0f);
    }
    e->pending_force = V(0.0f, 0.0f);

    /* contact solver: InitializeVelocityConstraints + WarmStart (dtRatio = 1) */
    for (int ci = 0; ci < nc; ++ci) {
        Contact* c = &con[ci];
        const int b = c->body;
        v2 xp; rot xq;
        body_xf(&B[b], b, &xp, &xq);
        v2 pts[2];
        world_manifold(&c->man, xp, xq, &c->normal, pts);
        c->vc_count = c->man.count;
        const float mB = im[b], iB = ii[b];
        for (int j = 0; j < c->man.count; ++j) {
            c->rB[j] = sub(pts[j], B[b].c);
            float rnB = cross(c->rB[j], c->normal);
            float kN = mB + iB * rnB * rnB;
            c->normal_mass[j] = kN > 0.0f ? 1.0f / kN : 0.0f;
            v2 tangent = cross_vs(c->normal, 1.0f);
            float rtB = cross(c->rB[j], tangent);
            float kT = mB + iB * rtB * rtB;
            c->tangent_mass[j] = kT > 0.0f ? 1.0f / kT : 0.0f;
            c->velocity_bias[j] = 0.0f;
            float vRel = dot(c->normal, add(B[b].v, cross_sv(B[b].w, c->rB[j])));
            if (vRel < -B2_VELOCITY_THRESHOLD) c->velocity_bias[j] = -0.0f * vRel; /* restitution 0 */
        }
        if (c->vc_count == 2) {
            float rn1B = cross(c->rB[0], c->normal), rn2B = cross(c->rB[1], c->normal);
            float k11 = mB + iB * rn1B * rn1B;
            float k22 = mB + iB * rn2B * rn2B;
            float k12 = mB + iB * rn1B * rn2B;
            if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
                c->K11 = k11; c->K12 = k12; c->K22 = k22;
                float det = k11 * k22 - k12 * k12;
                if (det != 0.0f) det = 1.0f / det;
                c->NM11 = det * k22; c->NM12 = -det * k12; c->NM21 = -det * k12; c->NM22 = det * k11;
            } else {
                c->vc_count = 1;
            }
        }
    }
    for (int ci = 0; ci < nc; ++ci) {
        Contact* c = &con[ci];
        const int b = c->body;
        v2 tangent = cross_vs(c->normal, 1.0f);
        for (int j = 0; j < c->vc_count; ++j) {
            v2 P = add(mul(c->nimp[j], c->normal), mul(c->timp[j], tangent));
            B[b].w += ii[b] * cross(c->rB[j], P);
            B[b].v = add(B[b].v, mul(im[b], P));
        }
    }

    /* joints: InitVelocityConstraints (+ warm start) — island order: joint 1 (i=+1) then joint 0 */
    v2 rA[2], rBj[2];
    float jm[2][9]; /* 3x3 mass matrix, column-major ex,ey,ez */
    float motor_mass[2];
    static const int JORDER[2] = {1, 0};
    for (int jo = 0; jo < 2; ++jo) {
        const int j = JORDER[jo];
        Joint* J = &e->joint[j];
        const int bA = 0, bB = 1 + j;
        rot qA = make_rot(B[bA].a), qB = make_rot(B[bB].a);
        rA[j] = rmul(qA, sub(V(0.0f, 0.0f), g_shapes.local_center[bA]));
        rBj[j] = rmul(qB, sub(joint_anchor_b(j), g_shapes.local_center[bB]));
        const float mA = im[bA], mB = im[bB], iA = ii[bA], iB = ii[bB];
        float* M = jm[j];
        M[0] = mA + mB + rA[j].y * rA[j].y * iA + rBj[j].y * rBj[j].y * iB; /* ex.x */
        M[3] = -rA[j].y * rA[j].x * iA - rBj[j].y * rBj[j].x * iB;          /* ey.x */
        M[6] = -rA[j].y * iA - rBj[j].y * iB;                               /* ez.x */
        M[1] = M[3];                                                         /* ex.y */
        M[4] = mA + mB + rA[j].x * rA[j].x * iA + rBj[j].x * rBj[j].x * iB; /* ey.y */
        M[7] = rA[j].x * iA + rBj[j].x * iB;                                /* ez.y */
        M[2] = M[6];                                                         /* ex.z */
        M[5] = M[7];                                                         /* ey.z */
        M[8] = iA + iB;                                                      /* ez.z */
        motor_mass[j] = iA + iB;
        if (motor_mass[j] > 0.0f) motor_mass[j] = 1.0f / motor_mass[j];
        {
            float jointAngle = B[bB].a - B[bA].a - joint_ref_angle(j);
            float lo = joint_lower(j), up = joint_upper(j);
            if (fabsf(up - lo) < 2.0f * B2_ANGULAR_SLOP) J->limit_state = 3;
            else if (jointAngle <= lo) { if (J->limit_state != 1) J->imp_z = 0.0f; J->limit_state = 1; }
            else if (jointAngle >= up) { if (J->limit_state != 2) J->imp_z = 0.0f; J->limit_state = 2; }
            else { J->limit_state = 0; J->imp_z = 0.0f; }
        }
        /* warm start, dtRatio = 1 */
        v2 P = V(J->imp_x, J->imp_y);
        B[bA].v = sub(B[bA].v, mul(mA, P));
        B[bA].w -= iA * (cross(rA[j], P) + J->motor_impulse + J->imp_z);
        B[bB].v = add(B[bB].v, mul(mB, P));
        B[bB].w += iB * (cross(rBj[j], P) + J->motor_impulse + J->imp_z);
    }

    /* velocity iterations */
    for (int it = 0; it < VEL_ITERS; ++it) {
        for (int jo = 0; jo < 2; ++jo) {
            const int j = JORDER[jo];
            Joint* J = &e->joint[j];
            const int bA = 0, bB = 1 + j;
            const float mA = im[bA], mB = im[bB], iA = ii[bA], iB = ii[bB];
            v2 vA = B[bA].v, vB = B[bB].v;
            float wA = B[bA].w, wB = B[bB].w;
            const float* M = jm[j];
            /* motor */
            if (J->limit_state != 3) {
                float Cdot = wB - wA - joint_motor_speed(j);
                float old = J->motor_impulse;
                float maxImp = h * (float)LEG_SPRING_TORQUE;
                J->motor_impulse = clampf(fmaf(-motor_mass[j], Cdot, old), -maxImp, maxImp);
                float impulse = J->motor_impulse - old;
                wA = fmaf(-iA, impulse, wA);
                wB = fmaf(iB, impulse, wB);
            }
            if (J->limit_state != 0) {
                v2 Cdot1 = sub_cross_sv(sub(add_cross_sv(vB, wB, rBj[j]), vA), wA, rA[j]);
                float Cdot2 = wB - wA;
                /* impulse = -Solve33(Cdot) */
                float ix, iy, iz;
                {
                    float exx = M[0], exy = M[1], exz = M[2], eyx = M[3], eyy = M[4], eyz = M[5], ezx = M[6], ezy = M[7], ezz = M[8];
                    /* det = dot(ex, cross(ey, ez)) */
                    float cx = eyy * ezz - eyz * ezy, cy = eyz * ezx - eyx * ezz, cz = eyx * ezy - eyy * ezx;
                    float det = exx * cx + exy * cy + exz * cz;
                    if (det != 0.0f) det = 1.0f / det;
                    float bx = Cdot1.x, by = Cdot1.y, bz = Cdot2;
                    /* x = det * dot(b, cross(ey, ez)) */
                    float sx = det * fmaf(bx, cx, fmaf(by, cy, bz * cz));
                    /* y = det * dot(ex, cross(b, ez)) */
                    float c2x = fmaf(by, ezz, -(bz * ezy)), c2y = fmaf(bz, ezx, -(bx * ezz)), c2z = fmaf(bx, ezy, -(by * ezx));
                    float sy = det * fmaf(exx, c2x, fmaf(exy, c2y, exz * c2z));
                    /* z = det * dot(ex, cross(ey, b)) */
                    float c3x = fmaf(eyy, bz, -(eyz * by)), c3y = fmaf(eyz, bx, -(eyx * bz)), c3z = fmaf(eyx, by, -(eyy * bx));
                    float sz = det * fmaf(exx, c3x, fmaf(exy, c3y, exz * c3z));
                    ix = -sx; iy = -sy; iz = -sz;
                }
                if (J->limit_state == 3) {
                    J->imp_x += ix; J->imp_y += iy; J->imp_z += iz;
                } else {
                    float newImpulse = J->imp_z + iz;
                    int violate = J->limit_state == 1 ? (newImpulse < 0.0f) : (newImpulse > 0.0f);
                    if (violate) {
                        v2 rhs = axpy(J->imp_z, V(M[6], M[7]), neg(Cdot1));
                        /* Solve22 */
                        float a11 = M[0], a12 = M[3], a21 = M[1], a22 = M[4];
                        float det = a11 * a22 - a12 * a21;
                        if (det != 0.0f) det = 1.0f / det;
                        float rx = det * fmaf(a22, rhs.x, -(a12 * rhs.y));
                        float ry = det * fmaf(a11, rhs.y, -(a21 * rhs.x));
                        ix = rx; iy = ry; iz = -J->imp_z;
                        J->imp_x += rx; J->imp_y += ry; J->imp_z = 0.0f;
                    } else {
                        J->imp_x += ix; J->imp_y += iy; J->imp_z += iz;
                    }
                }
                v2 P = V(ix, iy);
                vA = axpy(-mA, P, vA);
                wA = fmaf(-iA, fcross(rA[j], P) + iz, wA);
                vB = axpy(mB, P, vB);
                wB = fmaf(iB, fcross(rBj[j], P) + iz, wB);
            } else {
                v2 Cdot = sub_cross_sv(sub(add_cross_sv(vB, wB, rBj[j]), vA), wA, rA[j]);
                float a11 = M[0], a12 = M[3], a21 = M[1], a22 = M[4];
                float det = a11 * a22 - a12 * a21;
                if (det != 0.0f) det = 1.0f / det;
                float bx = -Cdot.x, by = -Cdot.y;
                v2 imp = V(det * fmaf(a22, bx, -(a12 * by)), det * fmaf(a11, by, -(a21 * bx)));
                J->imp_x += imp.x; J->imp_y += imp.y;
                vA = axpy(-mA, imp, vA);
                wA = fmaf(-iA, fcross(rA[j], imp), wA);
                vB = axpy(mB, imp, vB);
                wB = fmaf(iB, fcross(rBj[j], imp), wB);
            }
            B[bA].v = vA; B[bA].w = wA; B[bB].v = vB; B[bB].w = wB;
        }
        for (int ci = 0; ci < nc; ++ci) {
            Contact* c = &con[ci];
            const int b = c->body;
            const float mB = im[b], iB = ii[b];
            v2 vB = B[b].v;
            float wB = B[b].w;
            v2 normal = c->normal, tangent = cross_vs(normal, 1.0f);
            for (int j = 0; j < c->vc_count; ++j) {
                v2 dv = add_cross_sv(vB, wB, c->rB[j]);
                float vt = fdot(dv, tangent) - 0.0f;
                float maxF = c->friction * c->nimp[j];
                float newImp = clampf(fmaf(c->tangent_mass[j], -vt, c->timp[j]), -maxF, maxF);
                float lambda = newImp - c->timp[j];
                c->timp[j] = newImp;
                v2 P = mul(lambda, tangent);
                vB = axpy(mB, P, vB);
                wB = fmaf(iB, fcross(c->rB[j], P), wB);
            }
            if (c->vc_count == 1) {
                v2 dv = add_cross_sv(vB, wB, c->rB[0]);
                float vn = fdot(dv, normal);
                float newImp = fmaxf(fmaf(-c->normal_mass[0], vn - c->velocity_bias[0], c->nimp[0]), 0.0f);
                float lambda = newImp - c->nimp[0];
                c->nimp[0] = newImp;
                v2 P = mul(lambda, normal);
                vB = axpy(mB, P, vB);
                wB = fmaf(iB, fcross(c->rB[0], P), wB);
            } else {
                float a0 = c->nimp[0], a1 = c->nimp[1];
                v2 dv1 = add_cross_sv(vB, wB, c->rB[0]);
                v2 dv2 = add_cross_sv(vB, wB, c->rB[1]);
                float vn1 = fdot(dv1, normal), vn2 = fdot(dv2, normal);
                float bx = vn1 - c->velocity_bias[0], by = vn2 - c->velocity_bias[1];
                bx -= fmaf(c->K11, a0, c->K12 * a1);
                by -= fmaf(c->K12, a0, c->K22 * a1);
                float x0, x1;
                int solved = 0;
                /* case 1 */
                x0 = -fmaf(c->NM11, bx, c->NM21 * by);
                x1 = -fmaf(c->NM12, bx, c->NM22 * by);
                if (x0 >= 0.0f && x1 >= 0.0f) solved = 1;
                if (!solved) { /* case 2 */
                    x0 = -c->normal_mass[0] * bx; x1 = 0.0f;
                    vn2 = fmaf(c->K12, x0, by);
                    if (x0 >= 0.0f && vn2 >= 0.0f) solved = 1;
                }
                if (!solved) { /* case 3 */
                    x0 = 0.0f; x1 = -c->normal_mass[1] * by;
                    vn1 = fmaf(c->K12, x1, bx);
                    if (x1 >= 0.0f && vn1 >= 0.0f) solved = 1;
                }
                if (!solved) { /* case 4 */
                    x0 = 0.0f; x1 = 0.0f;
                    if (bx >= 0.0f && by >= 0.0f) solved = 1;
                }
                if (solved) {
                    float d0 = x0 - a0, d1 = x1 - a1;
                    v2 P1 = mul(d0, normal), P2 = mul(d1, normal);
                    vB = axpy(mB, add(P1, P2), vB);
                    wB = fmaf(iB, fcross(c->rB[0], P1) + fcross(c->rB[1], P2), wB);
                    c->nimp[0] = x0; c->nimp[1] = x1;
                }
            }
            B[b].v = vB; B[b].w = wB;
        }
    }

    /* StoreImpulses -> persistent slots */
    for (int ci = 0; ci < nc; ++ci) {
        ManifoldSlot* s = &new_slot[ci];
        s->key = con[ci].body * 16 + con[ci].edge;
        s->count = con[ci].man.count;
        for (int j = 0; j < con[ci].man.count; ++j) {
            s->id[j] = con[ci].man.id[j];
            /* points dropped by the block solver's redundancy test keep their warm-start value */
            s->nimp[j] = con[ci].nimp[j];
            s->timp[j] = con[ci].timp[j];
        }
    }
    memcpy(e->slot, new_slot, sizeof(new_slot));

    /* integrate positions */
    for (int b = 0; b < LL_NBODY; ++b) {
        v2 t = mul(h, B[b].v);
        if (dot(t, t) > B2_MAX_TRANSLATION * B2_MAX_TRANSLATION) {
            float ratio = B2_MAX_TRANSLATION / sqrtf(t.x * t.x + t.y * t.y);
            B[b].v = mul(ratio, B[b].v);
        }
        float rotn = h * B[b].w;
        if (rotn * rotn > B2_MAX_ROTATION * B2_MAX_ROTATION) {
            float ratio = B2_MAX_ROTATION / fabsf(rotn);
            B[b].w *= ratio;
        }
        B[b].c = add(B[b].c, mul(h, B[b].v));
        B[b].a += h * B[b].w;
    }

    /* position iterations */
    int position_solved = 0;
    for (int it = 0; it < POS_ITERS; ++it) {
        float min_sep = 0.0f;
        for (int ci = 0; ci < nc; ++ci) {
            Contact* c = &con[ci];
            const int b = c->body;
            const float mB = im[b], iB = ii[b];
            v2 cB = B[b].c;
            float aB = B[b].a;
            for (int j = 0; j < c->man.count; ++j) {
                rot qB = make_rot(aB);
                v2 pB = sub(cB, rmul(qB, g_shapes.local_center[b]));
                v2 normal, point;
                float separation;
                if (c->man.type == 0) {
                    normal = c->man.local_normal;
                    v2 plane = c->man.local_point;
                    v2 clip = add(rmul(qB, c->man.pt[j]), pB);
                    separation = dot(sub(clip, plane), normal) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
                    point = clip;
                } else {
                    v2 n = rmul(qB, c->man.local_normal);
                    v2 plane = add(rmul(qB, c->man.local_point), pB);
                    v2 clip = c->man.pt[j];
                    separation = dot(sub(clip, plane), n) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
                    point = clip;
                    normal = neg(n);
                }
                v2 rB = sub(point, cB);
                min_sep = fminf(min_sep, separation);
                float C = clampf(B2_BAUMGARTE * (separation + B2_LINEAR_SLOP), -B2_MAX_LINEAR_CORRECTION, 0.0f);
                float rnB = cross(rB, normal);
                float K = mB + iB * rnB * rnB;
                float impulse = K > 0.0f ? -C / K : 0.0f;
                v2 P = mul(impulse, normal);
                cB = add(cB, mul(mB, P));
                aB += iB * cross(rB, P);
            }
            B[b].c = cB; B[b].a = aB;
        }
        int contacts_ok = min_sep >= -3.0f * B2_LINEAR_SLOP;
        int joints_ok = 1;
        for (int jo = 0; jo < 2; ++jo) {
            const int j = JORDER[jo];
            Joint* J = &e->joint[j];
            const int bA = 0, bB = 1 + j;
            const float mA = im[bA], mB = im[bB], iA = ii[bA], iB = ii[bB];
            v2 cA = B[bA].c, cB = B[bB].c;
            float aA = B[bA].a, aB = B[bB].a;
            float angular_error = 0.0f, position_error;
            if (J->limit_state != 0) {
                float angle = aB - aA - joint_ref_angle(j);
                float limit_impulse = 0.0f;
                if (J->limit_state == 3) {
                    float C = clampf(angle - joint_lower(j), -B2_MAX_ANGULAR_CORRECTION, B2_MAX_ANGULAR_CORRECTION);
                    limit_impulse = -motor_mass[j] * C;
                    angular_error = fabsf(C);
                } else if (J->limit_state == 1) {
                    float C = angle - joint_lower(j);
                    angular_error = -C;
                    C = clampf(C + B2_ANGULAR_SLOP, -B2_MAX_ANGULAR_CORRECTION, 0.0f);
                    limit_impulse = -motor_mass[j] * C;
                } else {
                    float C = angle - joint_upper(j);
                    angular_error = C;
                    C = clampf(C - B2_ANGULAR_SLOP, 0.0f, B2_MAX_ANGULAR_CORRECTION);
                    limit_impulse = -motor_mass[j] * C;
                }
                aA -= iA * limit_impulse;
                aB += iB * limit_impulse;
            }
            {
                rot qA = make_rot(aA), qB = make_rot(aB);
                v2 ra = rmul(qA, sub(V(0.0f, 0.0f), g_shapes.local_center[bA]));
                v2 rb = rmul(qB, sub(joint_anchor_b(j), g_shapes.local_center[bB]));
                v2 C = sub(sub(add(cB, rb), cA), ra);
                position_error = sqrtf(C.x * C.x + C.y * C.y);
                float k11 = mA + mB + iA * ra.y * ra.y + iB * rb.y * rb.y;
                float k12 = -iA * ra.x * ra.y - iB * rb.x * rb.y;
                float k22 = mA + mB + iA * ra.x * ra.x + iB * rb.x * rb.x;
                float det = k11 * k22 - k12 * k12;
                if (det != 0.0f) det = 1.0f / det;
                v2 sol = V(det * (k22 * C.x - k12 * C.y), det * (k11 * C.y - k12 * C.x));
                v2 imp = neg(sol);
                cA = sub(cA, mul(mA, imp));
                aA -= iA * cross(ra, imp);
                cB = add(cB, mul(mB, imp));
                aB += iB * cross(rb, imp);
            }
            B[bA].c = cA; B[bA].a = aA; B[bB].c = cB; B[bB].a = aB;
            int ok = position_error <= B2_LINEAR_SLOP && angular_error <= B2_ANGULAR_SLOP;
            joints_ok = joints_ok && ok;
        }
        if (contacts_ok && joints_ok) { position_solved = 1; break; }
    }

    /* sleeping */
    {
        float min_sleep = 3.402823466e+38f;
        const float lin2 = B2_LINEAR_SLEEP_TOL * B2_LINEAR_SLEEP_TOL, ang2 = B2_ANGULAR_SLEEP_TOL * B2_ANGULAR_SLEEP_TOL;
        for (int b = 0; b < LL_NBODY; ++b) {
            if (B[b].w * B[b].w > ang2 || dot(B[b].v, B[b].v) > lin2) {
                B[b].sleep_time = 0.0f;
                min_sleep = 0.0f;
            } else {
                B[b].sleep_time += h;
                min_sleep = fminf(min_sleep, B[b].sleep_time);
            }
        }
        if (min_sleep >= B2_TIME_TO_SLEEP && position_solved) e->awake = 0;
    }
}

/* ---- public C API (ctypes) --------------------------------------------------------------------- */
LLEnv* ll_create(uint64_t seed, uint64_t env_id) {
    LLEnv* e = (LLEnv*)calloc(1, sizeof(LLEnv));
    e->seed = seed; e->env_id = env_id;
    shapes_init();
    return e;
}
void ll_destroy(LLEnv* e) { free(e); }

static void st_to_obs(const double st[8], float* obs) { for (int i = 0; i < 8; ++i) obs[i] = (float)st[i]; }

/* env.reset(): new world + the internal step(0) whose observation is returned */
void ll_reset(LLEnv* e, float* obs) {
    ll_begin_episode(e);
    double st[8];
    int term;
    (void)ll_env_step(e, 0, st, &term);
    st_to_obs(st, obs);
}

/* env.step() under TimeLimit(1000) with auto-reset: obs = first obs of the next episode when done */
void ll_step(LLEnv* e, int action, float* obs, float* next_obs, float* reward, uint8_t* terminated, uint8_t* truncated) {
    double st[8];
    int term;
    double r = ll_env_step(e, action, st, &term);
    e->elapsed += 1;
    int trunc = e->elapsed >= MAX_EPISODE_STEPS;
    e->ep_return += r;
    if (next_obs) st_to_obs(st, next_obs);
    *reward = (float)r;
    *terminated = (uint8_t)term;
    *truncated = (uint8_t)trunc;
    if (term || trunc) ll_reset(e, obs);
    else st_to_obs(st, obs);
}

/* gymnasium single-env semantics (oracle/gymnasium_shim): env.step() under TimeLimit(1000) WITHOUT auto-reset —
 * the caller calls ll_reset() itself after a done, exactly like the reference scripts do (ppo_lunarlander.py:220-223). */
void ll_step_single(LLEnv* e, int action, float* obs, float* reward, uint8_t* terminated, uint8_t* truncated) {
    double st[8];
    int term;
    double r = ll_env_step(e, action, st, &term);
    e->elapsed += 1;
    e->ep_return += r;
    st_to_obs(st, obs);
    *reward = (float)r;
    *terminated = (uint8_t)term;
    *truncated = (uint8_t)(e->elapsed >= MAX_EPISODE_STEPS);
}

void ll_get_state(const LLEnv* e, double* s) {
    int k = 0;
    for (int i = 0; i < CHUNKS; ++i) s[k++] = e->terrain_y[i];
    for (int b = 0; b < LL_NBODY; ++b) {
        s[k++] = e->body[b].c.x; s[k++] = e->body[b].c.y; s[k++] = e->body[b].a;
        s[k++] = e->body[b].v.x; s[k++] = e->body[b].v.y; s[k++] = e->body[b].w; s[k++] = e->body[b].sleep_time;
    }
    for (int j = 0; j < 2; ++j) {
        s[k++] = e->joint[j].imp_x; s[k++] = e->joint[j].imp_y; s[k++] = e->joint[j].imp_z;
        s[k++] = e->joint[j].motor_impulse; s[k++] = e->joint[j].limit_state;
    }
    s[k++] = e->pending_force.x; s[k++] = e->pending_force.y;
    s[k++] = e->game_over; s[k++] = e->leg_contact[0]; s[k++] = e->leg_contact[1]; s[k++] = e->awake;
    s[k++] = e->has_prev_shaping; s[k++] = e->prev_shaping;
    s[k++] = e->elapsed; s[k++] = e->episode; s[k++] = e->stepctr; s[k++] = e->ep_return;
    for (int i = 0; i < LL_MAX_MANIFOLDS; ++i) {
        const ManifoldSlot* m = &e->slot[i];
        s[k++] = m->key; s[k++] = m->count; s[k++] = m->id[0]; s[k++] = m->id[1];
        s[k++] = m->nimp[0]; s[k++] = m->nimp[1]; s[k++] = m->timp[0]; s[k++] = m->timp[1];
    }
    while (k < LL_STATE_DOUBLES) s[k++] = 0.0;
}

void ll_set_state(LLEnv* e, const double* s) {
    int k = 0;
    for (int i = 0; i < CHUNKS; ++i) e->terrain_y[i] = (float)s[k++];
    for (int b = 0; b < LL_NBODY; ++b) {
        e->body[b].c.x = (float)s[k++]; e->body[b].c.y = (float)s[k++]; e->body[b].a = (float)s[k++];
        e->body[b].v.x = (float)s[k++]; e->body[b].v.y = (float)s[k++]; e->body[b].w = (float)s[k++];
        e->body[b].sleep_time = (float)s[k++];
    }
    for (int j = 0; j < 2; ++j) {
        e->joint[j].imp_x = (float)s[k++]; e->joint[j].imp_y = (float)s[k++]; e->joint[j].imp_z = (float)s[k++];
        e->joint[j].motor_impulse = (float)s[k++]; e->joint[j].limit_state = (int)s[k++];
    }
    e->pending_force.x = (float)s[k++]; e->pending_force.y = (float)s[k++];
    e->game_over = (int)s[k++]; e->leg_contact[0] = (int)s[k++]; e->leg_contact[1] = (int)s[k++]; e->awake = (int)s[k++];
    e->has_prev_shaping = (int)s[k++]; e->prev_shaping = s[k++];
    e->elapsed = (int)s[k++]; e->episode = (uint32_t)s[k++]; e->stepctr = (uint32_t)s[k++]; e->ep_return = s[k++];
    for (int i = 0; i < LL_MAX_MANIFOLDS; ++i) {
        ManifoldSlot* m = &e->slot[i];
        m->key = (int)s[k++]; m->count = (int)s[k++]; m->id[0] = (uint32_t)s[k++]; m->id[1] = (uint32_t)s[k++];
        m->nimp[0] = (float)s[k++]; m->nimp[1] = (float)s[k++]; m->timp[0] = (float)s[k++]; m->timp[1] = (float)s[k++];
    }
}

int ll_state_doubles(void) { return LL_STATE_DOUBLES; }

/* body mass data, for the closed-form parity checks in tests/ */
void ll_mass_data(double* out /* [3][4]: inv_mass, inv_I, lcx, lcy */) {
    shapes_init();
    for (int b = 0; b < LL_NBODY; ++b) {
        out[4 * b + 0] = g_shapes.inv_mass[b]; out[4 * b + 1] = g_shapes.inv_I[b];
        out[4 * b + 2] = g_shapes.local_center[b].x; out[4 * b + 3] = g_shapes.local_center[b].y;
    }
}

/* Vectorised helpers for the CPU baseline timing: n independent envs stepped in a loop. */
void ll_vec_reset(LLEnv** envs, int n, float* obs) {
    for (int i = 0; i < n; ++i) ll_reset(envs[i], obs + 8 * i);
}
void ll_vec_step(LLEnv** envs, int n, const int32_t* actions, float* obs, float* next_obs, float* reward,
                 uint8_t* terminated, uint8_t* truncated) {
    for (int i = 0; i < n; ++i)
        ll_step(envs[i], actions[i], obs + 8 * i, next_obs ? next_obs + 8 * i : 0, reward + i, terminated + i, truncated + i);
}
