/* TEST INFRASTRUCTURE ONLY -- the oracle. Never linked, imported or executed by the product path
 * (binary_b200/, include/, libbinary_cuda.so). Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.
 *
 * Plain-C restatement of the reference's interval-overlap path (paths relative to /root/reference):
 *   - IntervalNode / BaseInterval          library/include/binary/algorithm/interval_tree.hpp:51-132
 *   - IntervalTree::insert_node_impl       interval_tree.hpp:230-260
 *   - RbTree::fix_insert / rotations       rb_tree.hpp:304-344, 255-302
 *   - IntervalTree::left/right_rotate      interval_tree.hpp:206-228 (max repair)
 *   - IntervalTree::find_overlaps_impl     interval_tree.hpp:306-328 (pruned PREORDER)
 *   - IntervalTree::find_overlap           interval_tree.hpp:290-304 (first hit)
 *   - one tree per group                   standalone/sv2nl/include/mapper.hpp:147-162
 * It reproduces the reference's tree SHAPE (root, colours, max) and native hit ORDER, so it can be
 * pinned against the reference's own known answers (test_interval_tree.cpp:87-155) and, in this
 * container, against oracle/_ref (the unmodified reference headers compiled by oracle/Makefile).
 * Parity status: PINNED (tests/test_oracle.py: golden vectors + _ref cross-check fixtures).
 *
 * The design differs from the reference on purpose (this is a restatement, not a copy): nodes live
 * in index-addressed arrays rather than unique_ptr-linked heap objects, and recursion is an
 * explicit stack.
 *
 * Also here: orc_brute (the bare predicate, interval_tree.hpp:119-121) and orc_flat_* (sort +
 * running max-end + scan, the CPU twin of the GPU index) used for full-size count/hash parity.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#define NIL (-1)
enum { RED = 0, BLACK = 1 };

typedef struct {
  uint32_t low, high, id; /* the interval + payload                       */
  uint32_t key, max;      /* key = low, max = subtree max of high  (:66-69) */
  int32_t left, right, parent;
  uint8_t color;
} orc_node;

typedef struct {
  orc_node *nodes; /* one pool for all groups */
  uint64_t n;
  uint32_t n_groups;
  uint32_t *group_val; /* sorted unique group values */
  int32_t *root;       /* per group root index       */
  uint64_t *group_size;
} orc_forest;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : (x > y);
}

static int find_group(const orc_forest *f, uint32_t g) {
  int lo = 0, hi = (int)f->n_groups - 1;
  while (lo <= hi) {
    int mid = (lo + hi) / 2;
    if (f->group_val[mid] == g) return mid;
    if (f->group_val[mid] < g) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

/* get_max(null) = numeric_limits<u32>::lowest() = 0            interval_tree.hpp:262-268 */
static inline uint32_t get_max(const orc_node *nd, int32_t i) { return i == NIL ? 0u : nd[i].max; }
static inline uint32_t max_u32(uint32_t a, uint32_t b) { return a > b ? a : b; }

/* recompute from interval.high and both children               interval_tree.hpp:275-278 */
static inline void refresh_max(orc_node *nd, int32_t i) {
  nd[i].max = max_u32(nd[i].high, max_u32(get_max(nd, nd[i].left), get_max(nd, nd[i].right)));
}

/* base rotation rb_tree.hpp:255-278, then the max repair of interval_tree.hpp:206-216:
 * the new subtree top inherits (at least) the old top's max, the demoted node is recomputed. */
static void rotate_left(orc_node *nd, int32_t *root, int32_t x) {
  int32_t y = nd[x].right;
  nd[x].right = nd[y].left;
  if (nd[x].right != NIL) nd[nd[x].right].parent = x;
  nd[y].parent = nd[x].parent;
  if (nd[x].parent == NIL) *root = y;
  else if (nd[nd[x].parent].left == x) nd[nd[x].parent].left = y;
  else nd[nd[x].parent].right = y;
  nd[y].left = x;
  nd[x].parent = y;
  nd[y].max = max_u32(nd[y].max, nd[x].max);
  refresh_max(nd, x);
}

static void rotate_right(orc_node *nd, int32_t *root, int32_t x) {
  int32_t y = nd[x].left;
  nd[x].left = nd[y].right;
  if (nd[x].left != NIL) nd[nd[x].left].parent = x;
  nd[y].parent = nd[x].parent;
  if (nd[x].parent == NIL) *root = y;
  else if (nd[nd[x].parent].left == x) nd[nd[x].parent].left = y;
  else nd[nd[x].parent].right = y;
  nd[y].right = x;
  nd[x].parent = y;
  nd[y].max = max_u32(nd[y].max, nd[x].max);
  refresh_max(nd, x);
}

static inline int is_red(const orc_node *nd, int32_t i) { return i != NIL && nd[i].color == RED; }

/* rb_tree.hpp:304-344 */
static void fix_insert(orc_node *nd, int32_t *root, int32_t z) {
  while (is_red(nd, nd[z].parent)) {
    int32_t p = nd[z].parent, g = nd[p].parent;
    if (p == nd[g].left) {
      int32_t u = nd[g].right;
      if (is_red(nd, u)) {
        nd[p].color = BLACK; nd[u].color = BLACK; nd[g].color = RED;
        z = g;
      } else {
        if (z == nd[p].right) { z = p; rotate_left(nd, root, z); }
        nd[nd[z].parent].color = BLACK;
        nd[nd[nd[z].parent].parent].color = RED;
        rotate_right(nd, root, nd[nd[z].parent].parent);
      }
    } else {
      int32_t u = nd[g].left;
      if (is_red(nd, u)) {
        nd[p].color = BLACK; nd[u].color = BLACK; nd[g].color = RED;
        z = g;
      } else {
        if (z == nd[p].left) { z = p; rotate_right(nd, root, z); }
        nd[nd[z].parent].color = BLACK;
        nd[nd[nd[z].parent].parent].color = RED;
        rotate_left(nd, root, nd[nd[z].parent].parent);
      }
    }
  }
  nd[*root].color = BLACK;
}

/* interval_tree.hpp:230-260: raise max on every node passed; strictly-less goes left, ties go right */
static void insert_node(orc_node *nd, int32_t *root, int32_t z) {
  int32_t x = *root, y = NIL;
  while (x != NIL) {
    y = x;
    nd[x].max = max_u32(nd[x].max, nd[z].max);
    x = (nd[z].key < nd[x].key) ? nd[x].left : nd[x].right;
  }
  nd[z].parent = y;
  if (y == NIL) *root = z;
  else if (nd[z].key < nd[y].key) nd[y].left = z;
  else nd[y].right = z;
  nd[z].color = RED;
  fix_insert(nd, root, z);
}

orc_forest *orc_build(uint64_t n, const uint32_t *group, const uint32_t *low, const uint32_t *high) {
  orc_forest *f = (orc_forest *)calloc(1, sizeof(orc_forest));
  if (!f) return NULL;
  f->n = n;
  f->nodes = (orc_node *)malloc((n ? n : 1) * sizeof(orc_node));
  /* unique group values */
  uint32_t *gv = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  uint64_t ng = 0;
  if (n) {
    if (group) { memcpy(gv, group, n * 4); qsort(gv, n, 4, cmp_u32); }
    else gv[0] = 0;
    uint64_t m = group ? n : 1;
    for (uint64_t i = 0; i < m; ++i) if (i == 0 || gv[i] != gv[i - 1]) gv[ng++] = gv[i];
  }
  f->n_groups = (uint32_t)ng;
  f->group_val = gv;
  f->root = (int32_t *)malloc((ng ? ng : 1) * sizeof(int32_t));
  f->group_size = (uint64_t *)calloc(ng ? ng : 1, sizeof(uint64_t));
  for (uint64_t g = 0; g < ng; ++g) f->root[g] = NIL;
  for (uint64_t i = 0; i < n; ++i) {
    orc_node *z = &f->nodes[i];
    z->low = low[i]; z->high = high[i]; z->id = (uint32_t)i;
    z->key = low[i]; z->max = high[i];
    z->left = z->right = z->parent = NIL; z->color = BLACK;
    int gi = find_group(f, group ? group[i] : 0u);
    insert_node(f->nodes, &f->root[gi], (int32_t)i);
    f->group_size[gi]++;
  }
  return f;
}

void orc_free(orc_forest *f) {
  if (!f) return;
  free(f->nodes); free(f->group_val); free(f->root); free(f->group_size); free(f);
}

uint64_t orc_size(const orc_forest *f, uint32_t group) {
  int gi = find_group(f, group);
  return gi < 0 ? 0 : f->group_size[gi];
}

int orc_root(const orc_forest *f, uint32_t group, uint32_t *key, uint32_t *low, uint32_t *high,
             uint32_t *id, uint32_t *max) {
  int gi = find_group(f, group);
  if (gi < 0 || f->root[gi] == NIL) return 0;
  const orc_node *r = &f->nodes[f->root[gi]];
  *key = r->key; *low = r->low; *high = r->high; *id = r->id; *max = r->max;
  return 1;
}

static int black_height(const orc_node *nd, int32_t i, int *ok, int *max_ok) {
  if (i == NIL) return 0;
  int l = black_height(nd, nd[i].left, ok, max_ok);
  int r = black_height(nd, nd[i].right, ok, max_ok);
  if (l != r) *ok = 0;
  uint32_t m = max_u32(nd[i].high, max_u32(get_max(nd, nd[i].left), get_max(nd, nd[i].right)));
  if (m != nd[i].max) *max_ok = 0;
  return l + (nd[i].color == BLACK);
}

/* black height, -1 on black-height mismatch (test_interval_tree.cpp:18-29), -2 on a wrong max */
int orc_check_invariants(const orc_forest *f, uint32_t group) {
  int gi = find_group(f, group);
  if (gi < 0) return 0;
  int ok = 1, max_ok = 1;
  int h = black_height(f->nodes, f->root[gi], &ok, &max_ok);
  if (!max_ok) return -2;
  return ok ? h : -1;
}

/* the predicate, evaluated on the query: low<=o.high && o.low<=high   interval_tree.hpp:119-121 */
static inline int overlaps(uint32_t ql, uint32_t qh, uint32_t tl, uint32_t th) {
  return ql <= th && tl <= qh;
}

/* interval_tree.hpp:290-304 */
int orc_find_overlap(const orc_forest *f, uint32_t group, uint32_t ql, uint32_t qh, uint32_t *low,
                     uint32_t *high, uint32_t *id) {
  int gi = find_group(f, group);
  if (gi < 0) return 0;
  const orc_node *nd = f->nodes;
  int32_t x = f->root[gi];
  while (x != NIL) {
    if (overlaps(ql, qh, nd[x].low, nd[x].high)) {
      *low = nd[x].low; *high = nd[x].high; *id = nd[x].id;
      return 1;
    }
    x = (ql <= get_max(nd, nd[x].left)) ? nd[x].left : nd[x].right;
  }
  return 0;
}

typedef struct { uint32_t *v; uint64_t n, cap; } u32vec;
static void push(u32vec *a, uint32_t x) {
  if (a->n == a->cap) {
    a->cap = a->cap ? a->cap * 2 : 1024;
    a->v = (uint32_t *)realloc(a->v, a->cap * 4);
  }
  a->v[a->n++] = x;
}

/* interval_tree.hpp:306-328 as an explicit-stack preorder: visit node, then left (if
 * q.low <= max(left)), then right (if q.high >= key && q.low <= max(right)). The right child is
 * pushed first so the left subtree is emitted first, which is the reference's native order. */
static uint64_t query_tree(const orc_node *nd, int32_t root, uint32_t ql, uint32_t qh, u32vec *out,
                           int32_t **stack, uint64_t *stack_cap) {
  uint64_t k = 0, sp = 0;
  if (root == NIL) return 0;
  if (*stack_cap < 256) { *stack_cap = 256; *stack = (int32_t *)realloc(*stack, 256 * 4); }
  (*stack)[sp++] = root;
  while (sp) {
    int32_t x = (*stack)[--sp];
    if (overlaps(ql, qh, nd[x].low, nd[x].high)) { push(out, nd[x].id); ++k; }
    if (sp + 2 > *stack_cap) { *stack_cap *= 2; *stack = (int32_t *)realloc(*stack, *stack_cap * 4); }
    int32_t l = nd[x].left, r = nd[x].right;
    if (r != NIL && qh >= nd[x].key && ql <= nd[r].max) (*stack)[sp++] = r;
    if (l != NIL && ql <= nd[l].max) (*stack)[sp++] = l;
  }
  return k;
}

typedef struct {
  const orc_forest *f;
  uint64_t b, e;
  const uint32_t *qg, *ql, *qh;
  uint64_t *counts;
  u32vec out;
} qjob;

static void *query_worker(void *arg) {
  qjob *j = (qjob *)arg;
  int32_t *stack = NULL; uint64_t cap = 0;
  for (uint64_t i = j->b; i < j->e; ++i) {
    int gi = find_group(j->f, j->qg ? j->qg[i] : 0u);
    j->counts[i] = gi < 0 ? 0
                          : query_tree(j->f->nodes, j->f->root[gi], j->ql[i], j->qh[i], &j->out,
                                       &stack, &cap);
  }
  free(stack);
  return NULL;
}

/* Batched find_overlaps. offsets: n_q+1 CSR; *targets: malloc'd ids in NATIVE preorder per query
 * (free with orc_free_buf); seconds: wall time of the query phase. Same contract as ref_query. */
int orc_query(const orc_forest *f, uint64_t n_q, const uint32_t *qg, const uint32_t *ql,
              const uint32_t *qh, int threads, uint64_t *offsets, uint32_t **targets,
              double *seconds) {
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n_q && n_q > 0) threads = (int)n_q;
  qjob *jobs = (qjob *)calloc(threads, sizeof(qjob));
  pthread_t *th = (pthread_t *)calloc(threads, sizeof(pthread_t));
  uint64_t *counts = (uint64_t *)malloc((n_q ? n_q : 1) * 8);
  double t0 = now_s();
  for (int t = 0; t < threads; ++t) {
    jobs[t].f = f; jobs[t].b = n_q * t / threads; jobs[t].e = n_q * (t + 1) / threads;
    jobs[t].qg = qg; jobs[t].ql = ql; jobs[t].qh = qh; jobs[t].counts = counts;
    if (threads > 1) pthread_create(&th[t], NULL, query_worker, &jobs[t]);
    else query_worker(&jobs[t]);
  }
  if (threads > 1) for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  if (seconds) *seconds = now_s() - t0;
  uint64_t acc = 0;
  for (uint64_t i = 0; i < n_q; ++i) { offsets[i] = acc; acc += counts[i]; }
  offsets[n_q] = acc;
  if (targets) {
    uint32_t *buf = (uint32_t *)malloc((acc ? acc : 1) * 4);
    uint64_t pos = 0;
    for (int t = 0; t < threads; ++t) {
      if (jobs[t].out.n) memcpy(buf + pos, jobs[t].out.v, jobs[t].out.n * 4);
      pos += jobs[t].out.n;
    }
    *targets = buf;
  }
  for (int t = 0; t < threads; ++t) free(jobs[t].out.v);
  free(jobs); free(th); free(counts);
  return 0;
}

void orc_free_buf(void *p) { free(p); }

/* The bare predicate over all pairs (same group only). ids ascending per query. O(n_t * n_q). */
int orc_brute(uint64_t n_t, const uint32_t *tg, const uint32_t *tl, const uint32_t *th_,
              uint64_t n_q, const uint32_t *qg, const uint32_t *ql, const uint32_t *qh,
              uint64_t *offsets, uint32_t **targets) {
  u32vec out = {0};
  for (uint64_t i = 0; i < n_q; ++i) {
    offsets[i] = out.n;
    uint32_t g = qg ? qg[i] : 0u;
    for (uint64_t t = 0; t < n_t; ++t)
      if ((tg ? tg[t] : 0u) == g && overlaps(ql[i], qh[i], tl[t], th_[t])) push(&out, (uint32_t)t);
  }
  offsets[n_q] = out.n;
  if (targets) { if (!out.v) out.v = (uint32_t *)malloc(4); *targets = out.v; } else free(out.v);
  return 0;
}

/* ---------------- flat index twin: sort by (group, low), running max of high, scan ------------- */

typedef struct { uint64_t key; uint32_t high, id; } flat_rec;
static int cmp_rec(const void *a, const void *b) {
  const flat_rec *x = (const flat_rec *)a, *y = (const flat_rec *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->id < y->id ? -1 : (x->id > y->id);
}

/* 64-bit finaliser (splitmix64's output function) used for the order-independent pair hash */
static inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

typedef struct {
  const flat_rec *rec; const uint32_t *runmax; uint64_t n_t;
  uint64_t b, e, qid_base;
  const uint32_t *qg, *ql, *qh;
  uint64_t *counts; /* optional, n_q */
  uint64_t total, hash;
} fjob;

static void *flat_worker(void *arg) {
  fjob *j = (fjob *)arg;
  uint64_t total = 0, hash = 0;
  for (uint64_t i = j->b; i < j->e; ++i) {
    uint64_t g = j->qg ? j->qg[i] : 0u;
    uint64_t gbeg_key = g << 32, ub_key = (g << 32) | j->qh[i];
    /* group start */
    uint64_t lo = 0, hi = j->n_t;
    while (lo < hi) { uint64_t m = (lo + hi) / 2; if (j->rec[m].key < gbeg_key) lo = m + 1; else hi = m; }
    uint64_t gbeg = lo;
    /* ub: first record with key > (g, q.high) */
    hi = j->n_t;
    while (lo < hi) { uint64_t m = (lo + hi) / 2; if (j->rec[m].key <= ub_key) lo = m + 1; else hi = m; }
    uint64_t ub = lo;
    /* lb: first record in [gbeg, ub) whose running max (per group) reaches q.low */
    lo = gbeg; hi = ub;
    while (lo < hi) { uint64_t m = (lo + hi) / 2; if (j->runmax[m] < j->ql[i]) lo = m + 1; else hi = m; }
    uint64_t k = 0;
    for (uint64_t t = lo; t < ub; ++t)
      if (j->rec[t].high >= j->ql[i]) {
        ++k;
        hash += mix64(((j->qid_base + i) << 32) | j->rec[t].id);
      }
    if (j->counts) j->counts[i] = k;
    total += k;
  }
  j->total = total; j->hash = hash;
  return NULL;
}

/* total hit count + order-independent hash sum(mix64(query_id<<32 | target_id)) mod 2^64.
 * qid_base is added to the local query index (for sharded checks). counts may be NULL. */
int orc_flat_count_hash(uint64_t n_t, const uint32_t *tg, const uint32_t *tl, const uint32_t *th_,
                        uint64_t n_q, const uint32_t *qg, const uint32_t *ql, const uint32_t *qh,
                        uint64_t qid_base, int threads, uint64_t *counts, uint64_t *total,
                        uint64_t *hash) {
  flat_rec *rec = (flat_rec *)malloc((n_t ? n_t : 1) * sizeof(flat_rec));
  uint32_t *runmax = (uint32_t *)malloc((n_t ? n_t : 1) * 4);
  for (uint64_t i = 0; i < n_t; ++i) {
    rec[i].key = ((uint64_t)(tg ? tg[i] : 0u) << 32) | tl[i];
    rec[i].high = th_[i]; rec[i].id = (uint32_t)i;
  }
  qsort(rec, n_t, sizeof(flat_rec), cmp_rec);
  for (uint64_t i = 0; i < n_t; ++i) {
    int fresh = (i == 0) || ((rec[i].key >> 32) != (rec[i - 1].key >> 32));
    runmax[i] = fresh ? rec[i].high : max_u32(runmax[i - 1], rec[i].high);
  }
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n_q && n_q > 0) threads = (int)n_q;
  fjob *jobs = (fjob *)calloc(threads, sizeof(fjob));
  pthread_t *th = (pthread_t *)calloc(threads, sizeof(pthread_t));
  for (int t = 0; t < threads; ++t) {
    jobs[t].rec = rec; jobs[t].runmax = runmax; jobs[t].n_t = n_t;
    jobs[t].b = n_q * t / threads; jobs[t].e = n_q * (t + 1) / threads; jobs[t].qid_base = qid_base;
    jobs[t].qg = qg; jobs[t].ql = ql; jobs[t].qh = qh; jobs[t].counts = counts;
    if (threads > 1) pthread_create(&th[t], NULL, flat_worker, &jobs[t]);
    else flat_worker(&jobs[t]);
  }
  uint64_t tot = 0, h = 0;
  for (int t = 0; t < threads; ++t) {
    if (threads > 1) pthread_join(th[t], NULL);
    tot += jobs[t].total; h += jobs[t].hash;
  }
  *total = tot; *hash = h;
  free(jobs); free(th); free(rec); free(runmax);
  return 0;
}

/* hash of an explicit pair list, same function as above (for checking GPU output on the host) */
uint64_t orc_pair_hash(uint64_t n, const uint32_t *query_id, const uint32_t *target_id) {
  uint64_t h = 0;
  for (uint64_t i = 0; i < n; ++i) h += mix64(((uint64_t)query_id[i] << 32) | target_id[i]);
  return h;
}

int orc_hardware_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
