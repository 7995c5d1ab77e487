/* putranse.h — C-ABI of libputranse.so, the B200-native drop-in for the hot path of
 * luofeisg/OpenKE-PuTransE (per-universe translational-embedding training + link prediction).
 *
 * The reference binds its native core with ctypes (openke/data/TrainDataLoader.py:30-31,
 * openke/data/TestDataLoader.py:30-31, openke/config/Tester.py:20-21: cdll.LoadLibrary("Base.so")).
 * This library is loaded the same way.  It exports
 *   (A) the Base.so symbols the reference's Python classes call on this path, with the same names,
 *       argument meaning and process-global state (section "reference-compatible surface"), and
 *   (B) new pk_* entry points that take DEVICE pointers (tensor.data_ptr()) and a CUDA stream and
 *       launch the hand-written sm_100a kernels.  All pk_* functions return 0 on success and a
 *       negative code on failure; pk_last_error() gives the text.  (The reference has no error
 *       convention: void returns, printf, exit(); reference SURVEY.md section 8(b).)
 *
 * Types follow reference openke/base/Setting.h:3-4: INT = long (int64), REAL = float.
 * No torch types appear anywhere in this file.  There is no CPU implementation of any pk_* compute
 * entry point: without a CUDA device they fail with PK_ERR_CUDA.
 */
#ifndef PUTRANSE_H
#define PUTRANSE_H
#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef long  PK_INT;   /* reference INT  */
typedef float PK_REAL;  /* reference REAL */

enum { PK_OK = 0, PK_ERR_ARG = -1, PK_ERR_STATE = -2, PK_ERR_CUDA = -3, PK_ERR_IO = -4, PK_ERR_UNSUPPORTED = -5 };
enum { PK_TRANSE = 0, PK_TRANSH = 1, PK_TRANSD = 2 };
enum { PK_SGD = 0, PK_ADAGRAD = 1 };

/* copies the calling thread's last error text into buf (NUL-terminated); returns its length */
int pk_last_error(char* buf, int n);
/* number of CUDA devices visible, or PK_ERR_CUDA */
int pk_cuda_device_count(void);
const char* pk_version(void);

/* ===================================================================================== (A)
 * reference-compatible surface (process-global state, single caller thread, caller-owned buffers)
 * ------------------------------------------------------------------------------------------- */
void   setInPath(char* path);                /* openke/base/Setting.h:12-19   */
void   setOutPath(char* path);               /* Setting.h:21-28               */
void   setWorkThreads(PK_INT threads);       /* Setting.h:36-39 : number of sampler streams */
PK_INT getWorkThreads(void);                 /* Setting.h:41-44               */
void   setBern(PK_INT con);                  /* Setting.h:92-95               */
PK_INT getEntityTotal(void);                 /* Setting.h:57-60  (universe-local after swapHelpers) */
PK_INT getRelationTotal(void);               /* Setting.h:62-65               */
PK_INT getTripleTotal(void);                 /* Setting.h:67-70               */
PK_INT getTrainTotal(void);                  /* Setting.h:72-75               */
PK_INT getTestTotal(void);                   /* Setting.h:77-80               */
PK_INT getValidTotal(void);                  /* Setting.h:82-85               */
void   setRandomSeed(PK_INT seed);           /* openke/base/Random.h:38-45 (srand)          */
PK_INT getRandomSeed(void);                  /* Random.h:47-50                */
void   randReset(void);                      /* Random.h:11-15 : seed the LCG streams from rand() */
void   importTrainFiles(void);               /* openke/base/Reader.h:169-234  */
void   importTestFiles(void);                /* Reader.h:246-342              */
/* openke/base/Base.cpp:266-310.  Buffers are HOST memory as in the reference; the batch is produced
 * by the CUDA sampler kernel (bit-identical to the reference's pthread sampler) and copied back.
 * Only mode 0, negRelRate 0, val_loss false are on the hot path; anything else is refused. */
void   sampling(PK_INT* batch_h, PK_INT* batch_t, PK_INT* batch_r, PK_REAL* batch_y, PK_INT batchSize,
                PK_INT negRate, PK_INT negRelRate, PK_INT mode, bool filter_flag, bool p, bool val_loss);
void   getParallelUniverse(PK_INT triple_constraint, PK_REAL balance_parameter); /* UniverseConstructor.h:327-397 */
PK_INT getEntityTotalUniverse(void);         /* openke/base/UniverseSetting.h:64-67 */
PK_INT getRelationTotalUniverse(void);       /* UniverseSetting.h:69-72       */
PK_INT getTrainTotalUniverse(void);          /* UniverseSetting.h:74-77       */
void   getEntityRemapping(PK_INT* ent_remapping);    /* UniverseSetting.h:79-84 */
void   getRelationRemapping(PK_INT* rel_remapping);  /* UniverseSetting.h:86-91 */
void   swapHelpers(void);                    /* UniverseSetting.h:123-154     */
void   resetUniverse(void);                  /* UniverseSetting.h:160-190     */
void   initTest(void);                       /* openke/base/Test.h:23-35      */
void   getHeadBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr);  /* Test.h:37-71   */
void   getTailBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr);  /* Test.h:73-107  */
void   validInit(void);                      /* openke/base/Valid.h:37-44     */
void   getValidHeadBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr); /* Valid.h:46-80  */
void   getValidTailBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr); /* Valid.h:82-114 */
/* Test.h:118-359 / Valid.h:116-240: `con` is a HOST score row in candidate order (slot 0 = truth).
 * The row is ranked by the CUDA ranking kernel; accumulators follow the reference's float sums. */
void   testHead(PK_REAL* con, PK_INT index, bool type_constrain);
void   testTail(PK_REAL* con, PK_INT index, bool type_constrain);
void   validHead(PK_REAL* con, PK_INT index);
void   validTail(PK_REAL* con, PK_INT index);
void   test_link_prediction(bool type_constrain);        /* Test.h:398-504  */
PK_REAL getTestLinkMRR(bool type_constrain);             /* Test.h:533-567  */
PK_REAL getTestLinkMR(bool type_constrain);
PK_REAL getTestLinkHit10(bool type_constrain);
PK_REAL getTestLinkHit3(bool type_constrain);
PK_REAL getTestLinkHit1(bool type_constrain);
PK_REAL getValidHit10(void);                             /* Valid.h:242-257 */
/* Adjacency queries over the id space the sampler currently reads (openke/base/Base.cpp:312-468; bound
 * unconditionally by the reference's TrainDataLoader, openke/data/TrainDataLoader.py:60-103).  Host work
 * over the sorted indexes.  getEntityRelations writes ALL distinct relations in order (the reference
 * never advances its output cursor, Base.cpp:441-468, and so leaves only the last one in slot 0). */
PK_INT getNumOfNegatives(PK_INT entity, PK_INT relation, bool entity_is_tail);
PK_INT getNumOfPositives(PK_INT entity, PK_INT relation, bool entity_is_tail);
void   getNegativeEntities(PK_INT* out, PK_INT entity, PK_INT relation, bool entity_is_tail);
void   getPositiveEntities(PK_INT* out, PK_INT entity, PK_INT relation, bool entity_is_tail);
PK_INT getNumOfEntityRelations(PK_INT entity, bool entity_is_tail);
void   getEntityRelations(PK_INT* out, PK_INT entity, bool entity_is_tail);
void   activateLoadOfAllTriples(bool unused);            /* openke/base/Reader.h:240-244 */
/* Triple classification inputs (openke/base/Test.h:573-599; bound by TestDataLoader.py:49-56): every
 * test triple plus one corrupted twin.  Twin i consumes two draws of LCG stream 0 (coin, entity) and is
 * corrupted with the FILTERED corruptors against the training index, so the device kernel jumps the
 * stream ahead by 2i and is bit-identical to the reference's sequential loop.  HOST buffers [testTotal]. */
void   getNegTest(void);
void   getTestBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr, PK_INT* nh, PK_INT* nt, PK_INT* nr);

/* Incremental setting (openke/base/Incremental.h; bound by openke/data/IncrementalTrainDataLoader.py:40-60 and
 * IncrementalTestDataLoader.py:34-66).  <inPath>/incremental/{entity2id,relation2id}.txt give the global id space,
 * <inPath>/incremental/<s>/train-op2id.txt the "h t r +|-" operations of snapshot s, .../global_triple2id.txt the
 * snapshot's whole triple list (filter set + candidate entities), .../{test,valid}2id.txt its evaluation lists. */
void   activateIncrementalSetting(void);          /* Incremental.h:45-48   */
void   initializeIncrementalSetting(void);        /* :207-217              */
void   setNumSnapshots(PK_INT n);                 /* :52-60                */
PK_INT getNumSnapshots(void);
void   setNumOperationsRate(PK_INT n);            /* :141-144              */
void   readGlobalNumEntities(void);               /* :182-193              */
void   readGlobalNumRelations(void);              /* :195-205              */
void   initializeTrainingOperations(int snapshot);/* :299-321              */
void   evolveTrainList(void);                     /* :798-846 : replay the operations, rebuild every training index */
void   loadSnapshotTriples(int snapshot);         /* :891-924              */
void   loadTestData(int snapshot);                /* :248-271              */
void   loadValidData(int snapshot);               /* :273-296              */
PK_INT getNumCurrentlyContainedEntities(void);    /* :134-138              */
void   initializeTripleOperations(int snapshot);  /* :323-348 : bound by IncrementalTestDataLoader.py:38; refuses at call time */
void   evolveTripleList(void);                    /* :926-949 : superseded by loadSnapshotTriples; refuses at call time */
int    pk_incremental_reset(void);
int64_t pk_incremental_list(int which, int32_t* out);

/* ===================================================================================== (B)
 * host-side exports of the graph state (for uploading to the device); ids are int32, triples are
 * (h, r, t) records of 3 x int32
 * ------------------------------------------------------------------------------------------- */
int pk_import_count(void);                       /* how often importTrainFiles ran (Bernoulli drift) */
/* current sampler id space (global graph, or the universe after swapHelpers) */
int pk_train_index(int32_t* by_head /*[nT*3] sorted (h,r,t)*/, int32_t* by_tail /*[nT*3] sorted (t,r,h)*/,
                   float* left_mean /*[nR]*/, float* right_mean /*[nR]*/);
int pk_get_lcg(uint64_t* state, int n /* capacity of state; min(n, workThreads) streams are copied */);
int pk_set_lcg(const uint64_t* state, int n);
/* acc = (float)(acc + v[i]) in order: how the reference's float metric accumulators sum doubles (Test.h:213-223) */
float pk_f32_running_sum(const double* v, int64_t n);
int pk_universe_triples(int32_t* collected_global /*[nT*3] collection order*/);
/* which: 0 test, 1 valid.  triples sorted (r,h,t) as the reference's testList/validList */
int pk_eval_triples(int which, int32_t* hrt /*[n*3]*/);
/* side: 0 head prediction, 1 tail prediction.  CSR of known-true candidates per query (truth
 * excluded, ascending).  Pass cand = NULL to get the total count in *n_cand first. */
int pk_filter_csr(int which, int side, int64_t* offsets /*[n+1]*/, int32_t* cand, int64_t* n_cand);

/* ---- many universes at once, built on nthreads host threads (re-entrant; global graph read-only).
 * Universe i is srand(seeds[i]); randReset(); getParallelUniverse(tcs[i], balances[i]). */
typedef struct pk_universe_set pk_universe_set;
pk_universe_set* pk_universes_build(int n, const int64_t* seeds, const int64_t* tcs, const float* balances,
                                    int nthreads);
/* without the (t,r,h) order, per-entity ranges and Bernoulli means (enough for filter_flag = 0, bern_flag = 0) */
pk_universe_set* pk_universes_build_lean(int n, const int64_t* seeds, const int64_t* tcs, const float* balances,
                                         int nthreads);
void pk_universes_free(pk_universe_set* s);
int  pk_universes_count(const pk_universe_set* s);
int  pk_universes_sizes(const pk_universe_set* s, int64_t* n_tri, int64_t* n_ent, int64_t* n_rel, int64_t* focus);
/* packed exports; universe i occupies [prefix(n_tri)[i], ...) etc.  Any pointer may be NULL. */
int  pk_universes_export(const pk_universe_set* s, int32_t* tri_by_head, int32_t* tri_by_tail,
                         int32_t* tri_collected_global, int32_t* ent_remap, int32_t* rel_remap,
                         float* left_mean, float* right_mean, uint64_t* lcg /*[n*workThreads]*/);

/* Hyper-parameters of n universes as the reference draws them from Python's `random` after random.seed(seeds[i])
 * (openke/config/Parallel_Universe_Config.py:157-161,211-212,232,237-240): tc = randrange(tc_lo, tc_hi); balance =
 * round(uniform(bal_lo, bal_hi), 2); margin = randrange(..); epochs = randrange(..) if draw_epochs else epochs_lo;
 * lr = round(uniform(lr_lo, lr_hi), lr_digits).  CPython's generator restated; PK_ERR_UNSUPPORTED for ranges it does not
 * cover (the caller then uses random.Random). */
int pk_python_hyper_draws(int n, const int64_t* seeds, int64_t tc_lo, int64_t tc_hi, double bal_lo, double bal_hi,
                          int64_t margin_lo, int64_t margin_hi, int64_t epochs_lo, int64_t epochs_hi, int draw_epochs,
                          double lr_lo, double lr_hi, int lr_digits, int64_t* tc, double* balance, int64_t* margin,
                          int64_t* epochs, double* lr);

/* ---- the same universes built ON THE GPU, one warp per universe (csrc/walk_device.cu); replaces the host threads of
 * pk_universes_build_lean for reference openke/base/UniverseConstructor.h:39-67,92-233,327-397 (getParallelUniverse,
 * get_entity_subset, the bidirectional walk, enumerateTrainListUniverse) after Random.h:11-15,38-45 (srand + randReset).
 * Bit-identical to the host builder (and through it to the reference).  Lean universes only: the local (h,r,t) list
 * and the two remaps, which is what training with filter_flag = 0 and bern_flag = 0 reads.  Fixed strides per universe:
 * PK_WALK_CAP triples, 2*PK_WALK_CAP entities, PK_WALK_CAP relations.  A universe the kernel does not handle reports a
 * status != PK_WALK_OK in its sizes row and must be built by pk_universes_build_lean. */
#define PK_WALK_CAP 2048
enum { PK_WALK_OK = 0, PK_WALK_TOO_LARGE = 1 /* tc or starting points > PK_WALK_CAP */, PK_WALK_ISOLATED = 2 /* entity without triples */,
       PK_WALK_EMPTY = 3 /* the walk collected nothing (the host builder reports the error) */ };
int pk_walk_cap(void);
int pk_walk_device_check(void);                      /* 0 = the current training graph can be walked on the device */
int pk_walk_scratch_bytes(int n, int64_t* out3);     /* bytes of d_bitmaps, d_got, d_trees for n universes */
/* h_lcg [n*workThreads] and h_focus [n] are HOST outputs, filled before the call returns; the d_* outputs are valid when
 * `stream` reaches the end of the launch: d_tri [n][CAP][3] local ids sorted (h,r,t), d_ent_remap [n][2*CAP],
 * d_rel_remap [n][CAP], d_sizes [n][8] = nT nE nR focus draws status rounds 0, d_got [n][CAP][3] collected triples in
 * global ids (collection order).  d_bitmaps / d_trees are scratch (d_trees may be NULL when its size is 0). */
int pk_universes_walk_device(int n, const int64_t* seeds, const int64_t* tcs, const float* balances, uint64_t* h_lcg,
                             int64_t* h_focus, uint32_t* d_bitmaps, int32_t* d_got, uint32_t* d_trees, int32_t* d_tri,
                             int32_t* d_ent_remap, int32_t* d_rel_remap, int32_t* d_sizes, void* stream);
/* the strided remaps of universes 0..n-1 back to back: d_ent_packed [sum nE], d_rel_packed [sum nR] (what the
 * evaluation indexes; the caller knows the sums from the sizes rows) */
int pk_walk_pack_remaps(int n, const int32_t* d_sizes, const int32_t* d_ent_remap, const int32_t* d_rel_remap,
                        int32_t* d_ent_packed, int32_t* d_rel_packed, void* stream);

/* Initial tables of n embedding spaces on host threads, bit-identical to the reference's model
 * constructors after torch.manual_seed(seeds[i]) (reference openke/module/model/TransE.py:17-22,
 * TransH.py:17-24, TransD.py:18-27; Parallel_Universe_Config.py:157-161): every table first
 * consumes torch's default nn.Embedding normal_() draw, then all tables are drawn
 * uniform(-bounds, +bounds) from the same mt19937 stream, in table order.
 * rows/row_off/bounds are [n * n_tables] (space-major); out[t] is the packed [sum rows, dims[t]]
 * HOST buffer of table t.  fused = 1 evaluates x*(to-from)+from with one rounding (how torch's
 * AVX2/AVX-512 builds contract it), 0 with two. */
int pk_torch_init_tables(int n, const int64_t* seeds, int n_tables, const int64_t* rows, const int32_t* dims,
                         float* const* out, const int64_t* row_off, const double* bounds, int fused, int nthreads);

/* ===================================================================================== (C)
 * CUDA entry points.  Pointers named d_* are device pointers; `stream` is a cudaStream_t.
 * ------------------------------------------------------------------------------------------- */

/* Shape of one embedding space and of one optimisation problem on it. */
typedef struct {
    int32_t model;       /* PK_TRANSE / PK_TRANSH / PK_TRANSD                                   */
    int32_t dim;         /* TransD: dim_e == dim_r == dim (the only case the reference runs)    */
    int32_t p_norm;      /* 1 or 2                                                              */
    int32_t norm_flag;   /* reference TransE.py:47-50                                           */
    int32_t opt;         /* PK_SGD / PK_ADAGRAD (reference Trainer.py:65-88)                    */
    int32_t neg_ent;     /* k negatives per positive (reference TrainDataLoader neg_ent)        */
    int32_t bern;        /* reference setBern                                                   */
    int32_t filter;      /* reference filter_flag                                               */
    int32_t work_threads;/* number of LCG streams the batch is sliced over (reference threads)  */
    int32_t reserved;
} pk_model_cfg;

/* Device-resident parameter tables of one embedding space (row-major [rows, dim] fp32).
 * ent[0]=ent_embeddings, ent[1]=ent_transfer (TransD);  rel[0]=rel_embeddings,
 * rel[1]=norm_vector (TransH) or rel_transfer (TransD).  *_state = Adagrad sum of squares. */
typedef struct {
    float* ent[2];
    float* rel[2];
    float* ent_state[2];
    float* rel_state[2];
    int64_t n_ent, n_rel;
} pk_tables;

/* Device-resident sampler index of one id space. */
typedef struct {
    const int32_t* by_head;    /* [n_tri*3] (h,r,t) sorted (h,r,t)  == reference trainList/trainHead */
    const int32_t* by_tail;    /* [n_tri*3] (h,r,t) sorted (t,r,h)  == reference trainTail; NULL if !filter */
    const float*   left_mean;  /* [n_rel] ; NULL if !bern */
    const float*   right_mean;
    uint64_t*      lcg;        /* [work_threads] stream states, advanced in place */
    int64_t n_tri, n_ent, n_rel;
    /* optional: [n_ent + 1] first record of every head in by_head / of every tail in by_tail.  They only
     * shorten the binary searches of filtered corruption (Corrupt.h:27-56) from the whole index to one
     * entity's records; NULL = search the whole index.  Results are identical either way. */
    const int64_t* head_off;
    const int64_t* tail_off;
} pk_sampler;

/* pk_torch_init_tables on the DEVICE: same arguments (host arrays), but d_out[t] are device tables.
 * One thread block per space replays torch's MT19937 stream; removes the host RNG time and the H2D
 * copy of the tables from the end-to-end path.  Bit-identical to the host version (tests). */
int pk_init_tables_device(int n, const int64_t* seeds, int n_tables, const int64_t* rows, const int32_t* dims,
                          float* const* d_out, const int64_t* row_off, const double* bounds, int fused, void* stream);

/* K0: one reference sampling() call on the device.  d_h/d_t/d_r are int32 [B*(1+k)] in the
 * reference's layout [B positives | B negatives#1 | ...] (Base.cpp:216-232). */
int pk_sample_batch(const pk_model_cfg* cfg, const pk_sampler* smp, int64_t batch_size,
                    int32_t* d_h, int32_t* d_t, int32_t* d_r, void* stream);

/* Workspace for the single-space train step (K1); opaque, device-resident. */
typedef struct pk_workspace pk_workspace;
pk_workspace* pk_workspace_create(const pk_model_cfg* cfg, int64_t n_ent, int64_t n_rel, int64_t max_batch);
void pk_workspace_free(pk_workspace* ws);
/* synchronises `stream` and reports whether a train step refused its batch: an id out of range, a
 * negative whose relation differs from its positive's (relation corruption is not on this path), or
 * a negative that replaces BOTH entities of its positive (the reference sampler replaces one,
 * Base.cpp:216-232).  A refused batch leaves the tables untouched and its loss is NaN. */
int pk_workspace_check(pk_workspace* ws, void* stream);

/* K1: one fused train step on given ids: gather, project, normalise, energy, margin loss, analytic
 * backward, per-row gradient accumulation, SGD/Adagrad scatter into the tables in place.
 * Equivalent of Trainer.train_one_step (reference Trainer.py:44-56).  d_loss[0] receives the loss. */
int pk_train_step(const pk_model_cfg* cfg, const pk_tables* tab, pk_workspace* ws, int64_t batch_size,
                  const int32_t* d_h, const int32_t* d_t, const int32_t* d_r, float margin, float lr,
                  float* d_loss, void* stream);

/* K0+K1 looped: `steps` consecutive sampling()+train_one_step pairs without leaving the device.
 * 64 steps are captured into a CUDA graph that the workspace keeps and re-launches for later calls
 * with identical arguments (an epoch loop); the call returns without synchronising.  d_loss is
 * [steps]; smp->lcg ends `steps` batches further, exactly like `steps` reference sampling() calls.
 * Equivalent of the body of Trainer.run (reference Trainer.py:91-99). */
int pk_train_steps(const pk_model_cfg* cfg, const pk_tables* tab, const pk_sampler* smp, pk_workspace* ws,
                   int64_t batch_size, int64_t steps, float margin, float lr, float* d_loss, void* stream);

/* K2: many universes, one launch.  One thread block per universe runs all epochs x nbatches steps
 * of that universe with its tables staged in shared memory when they fit.  The library keeps a small
 * device buffer per calling thread for the descriptors and for the gradient sums of entity rows that
 * occur more than once in a batch ((2 + neg_ent) * max batch_size rows per universe, L2-resident);
 * everything else is caller-owned.  Entity rows that occur several times in one batch are summed with
 * floating-point reductions whose order is not fixed: results are reproducible to rounding, not bit
 * for bit (the reference's autograd scatter-add on a GPU has the same property). */
typedef struct {
    int64_t tri_off;   /* first record of this universe in the packed by_head / by_tail arrays   */
    int64_t ent_off;   /* first row in the packed entity tables / entity state                  */
    int64_t rel_off;   /* first row in the packed relation tables                               */
    int64_t loss_off;  /* first slot in d_loss (one float per step), or -1                      */
    int32_t n_tri, n_ent, n_rel;
    int32_t batch_size, nbatches, epochs;
    float   margin, lr;
    uint64_t lcg[8];   /* stream states (work_threads <= 8 on this path)                        */
} pk_universe_desc;

int pk_train_universes(const pk_model_cfg* cfg, const pk_tables* packed, const int32_t* d_by_head,
                       const int32_t* d_by_tail, const float* d_left_mean, const float* d_right_mean,
                       const pk_universe_desc* h_desc /*HOST array*/, int n_universes, float* d_loss,
                       void* stream);
/* How pk_train_universes would run a universe of this shape: 0 = tables staged in shared memory,
 * 1 = entity tables in global memory (relation tables and batch scratch in shared memory),
 * 2 = does not fit the universe kernel (relation-rich graph or very large batch): train it with
 * pk_train_steps on its slice of the packed tables.  Negative on error. */
int pk_universe_kernel_class(const pk_model_cfg* cfg, int64_t n_ent, int64_t n_rel, int64_t batch_size);
/* Profiling aid: when d_buf is non-NULL every universe block of later pk_train_universes calls
 * writes (loss_off, nanoseconds it ran) into d_buf[2i], d_buf[2i+1] (device, 2*n_universes int64). */
int pk_debug_universe_timer(long long* d_buf);
/* number of kernel launches the last pk_* call on this thread issued (for bench accounting) */
int pk_last_launch_count(void);

/* K3: link prediction of one embedding space over n test triples, both sides, raw + filtered.
 * d_triples int32 [n*3] (h,r,t); filter CSR per side as produced by pk_filter_csr (device copies).
 * d_ranks int32 [n*4] = head raw, head filtered, tail raw, tail filtered (0-based count of better
 * candidates, reference Test.h:159-167). */
int pk_rank_space(const pk_model_cfg* cfg, const pk_tables* tab, int64_t n, const int32_t* d_triples,
                  const int64_t* d_foff_head, const int32_t* d_fcand_head, const int64_t* d_foff_tail,
                  const int32_t* d_fcand_tail, int32_t* d_ranks, void* stream);

/* K3u: PuTransE global energy estimation.  For every work item (key row, universe, local fixed
 * entity, local relation, side) scores all local entities of that universe and folds the result
 * into d_energy[key_row, global entity] with an elementwise minimum
 * (reference Parallel_Universe_Config.py:446-465,516-543). d_energy is [n_keys, n_ent_global] fp32,
 * +inf where no universe speaks.  After the per-GPU pass the caller may min-all-reduce d_energy
 * across ranks (NCCL, op=min) before ranking. */
typedef struct {
    int32_t key_row, universe, fixed_local, rel_local, side /*0 head batch, 1 tail batch*/, reserved;
} pk_energy_item;
int pk_universe_energies(const pk_model_cfg* cfg, const pk_tables* packed, const int64_t* d_ent_off,
                         const int64_t* d_rel_off, const int32_t* d_n_ent, const int32_t* d_ent_remap,
                         const pk_energy_item* d_items, int64_t n_items, float* d_energy,
                         int64_t n_ent_global, void* stream);
/* PuTransE missing_embedding_handling='null_vector' (reference Parallel_Universe_Config.py:378-388,
 * 494-514,634-640).  pk_universe_tuple_scores folds, for every work item, the score of (zero vector,
 * relation, fixed entity) of that universe into d_tuple[key_row] with a minimum (d_tuple must start
 * at +inf; min-all-reduce it across ranks like the energies); pk_fill_missing_energies then gives
 * every still-+inf candidate of a key row that row's tuple score, if it is finite. */
int pk_universe_tuple_scores(const pk_model_cfg* cfg, const pk_tables* packed, const int64_t* d_ent_off,
                             const int64_t* d_rel_off, const pk_energy_item* d_items, int64_t n_items, float* d_tuple,
                             void* stream);
int pk_fill_missing_energies(float* d_energy, int64_t n_rows, int64_t n_ent_global, const float* d_tuple, void* stream);
/* rank n queries from energy rows: query i reads row d_key_row[i], truth entity d_truth[i];
 * known-true candidates from the CSR.  Implements Test.h:118-238 including the +inf branch
 * (:181-206).  d_ranks int32 [n*2] = raw, filtered. */
int pk_rank_from_energy(const float* d_energy, int64_t n_ent_global, int64_t n, const int32_t* d_key_row,
                        const int32_t* d_truth, const int64_t* d_foff, const int32_t* d_fcand,
                        int32_t* d_ranks, void* stream);
/* one HOST-ordered row in the reference's candidate order (slot 0 = truth, then every other entity
 * ascending; Test.h:61-68): the kernel behind testHead/testTail.  d_truth1: int32[1]; d_foff2:
 * int64[2] = {0, #known}; d_ranks2: int32[2] = raw, filtered. */
/* pk_rank_from_energy restricted to the candidates mask[e] != 0 (incremental setting: the entities the snapshot
 * currently contains, Test.h:181-206); an unscored truth ranks n_candidates */
int pk_rank_from_energy_masked(const float* d_energy, int64_t n_ent_global, int64_t n, const int32_t* d_key_row,
                               const int32_t* d_truth, const int64_t* d_foff, const int32_t* d_fcand, int32_t* d_ranks,
                               const uint8_t* d_candidate_mask, int64_t n_candidates, void* stream);
int pk_rank_candidate_row(const float* d_con, int64_t n_ent, const int32_t* d_truth1, const int64_t* d_foff2,
                          const int32_t* d_fcand, int32_t* d_ranks2, void* stream);
/* Model.forward / Model.predict on int64 index batches (reference TransE.py:62-74,88-94 and the
 * TransH/TransD twins).  Output length n = max(nh, nt, nr); arrays shorter than n broadcast by
 * modulo, which is what the reference's view(-1, r.shape[0], dim) does.  head_batch = 1 computes
 * h + (r - t), otherwise (h + r) - t.  *d_bad_flag is set to 1 if an index is out of range. */
int pk_score_batch(const pk_model_cfg* cfg, const pk_tables* tab, const int64_t* d_h, int64_t nh, const int64_t* d_t,
                   int64_t nt, const int64_t* d_r, int64_t nr, int head_batch, float* d_out, int* d_bad_flag,
                   void* stream);
/* device self-test: compares the train kernels' division / square-root helpers bit-for-bit with
 * IEEE __fdiv_rn / __fsqrt_rn on n pseudo-random operand pairs.  mismatches3 = {division, sqrt,
 * exact-zero cases} (HOST int64[3]); all three must be 0. */
int pk_selftest_arith(int64_t n, uint32_t seed, int64_t* mismatches3);
/* fill with +inf */
int pk_fill_inf(float* d, int64_t n, void* stream);
/* d_out[r] = d_in[r] / max(||d_in[r]||_2, eps), the ranking kernels' own normalisation (reference TransE.py:52-55
 * F.normalize): PuTransE evaluation normalises a TransE universe's tables once instead of once per (key, universe) */
int pk_normalise_rows(const float* d_in, float* d_out, int64_t rows, int d, void* stream);

#ifdef __cplusplus
}
#endif
#endif
