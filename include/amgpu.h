/* amgpu — C ABI of the B200-native bulk change-replay engine for automerge-classic.
 *
 * This is the drop-in boundary: a backend module for `Automerge.setDefaultBackend()`
 * (reference src/automerge.js:147-149; function set in backend/index.js:1-8 and
 * @types/automerge/index.d.ts:139-162) binds exactly these entry points through N-API (see
 * INTEGRATION.md for the shim); the Python mirror in automerge_classic_b200/engine.py binds them
 * through ctypes. Plain pointers and sizes only. Every call is synchronous (the reference backend is
 * synchronous: backend/backend.js:27-32). There is no CPU fallback: amg_init fails when no CUDA
 * device is present.
 *
 * Error convention (reference: synchronous `throw` of RangeError / TypeError, state unchanged,
 * backend/new.js:1793-1795): functions return 0 on success, otherwise an amg_error_code, and fill
 * `err->msg` with the reference's message text (e.g. "no matching operation for pred: 3@abcd").
 */
#ifndef AMGPU_H
#define AMGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct amg_backend amg_backend;   /* one document: replaces class BackendDoc, backend/new.js:1694 */
typedef struct amg_patch amg_patch;       /* flat binary patch, see "patch layout" below */
typedef struct amg_buffers amg_buffers;   /* a list of byte buffers returned by the library */

typedef enum { AMG_OK = 0, AMG_RANGE_ERROR = 1, AMG_TYPE_ERROR = 2, AMG_INTERNAL_ERROR = 3, AMG_UNSUPPORTED = 4, AMG_CUDA_ERROR = 5 } amg_error_code;
typedef struct { int code; char msg[512]; } amg_error;

/* Backend.init()  — backend/backend.js:8-10 (new BackendDoc(), new.js:1751-1767) */
amg_backend* amg_init(int cuda_device, amg_error* err);
/* Backend.clone() — backend/backend.js:12-14 (BackendDoc.clone, new.js:1773-1790) */
amg_backend* amg_clone(amg_backend* b, amg_error* err);
/* Backend.load(data) — backend/backend.js:104-107 -> new BackendDoc(buffer), new.js:1709-1750 (document chunk, type 0) */
amg_backend* amg_load(int cuda_device, const uint8_t* data, size_t len, amg_error* err);
/* Backend.free()  — backend/backend.js:16-19 */
void amg_free(amg_backend* b);

/* Drops the document but keeps every device / pinned allocation (steady-state serving, benchmarking). */
int amg_reset(amg_backend* b, amg_error* err);
/* Pre-sizes the change arena (pinned host mirror + device) so that a bulk replay does not grow it mid-call. */
int amg_reserve(amg_backend* b, size_t arena_bytes, amg_error* err);

/* Backend.applyChanges(state, changes) — backend/backend.js:27-32 -> BackendDoc.applyChanges, new.js:1797-1879.
 * `bufs[i]` / `lens[i]`: the binary changes (chunk type 1, or 2 = DEFLATE, inflated on the device by an RFC 1951 decoder,
 * csrc/inflate.cuh, where columnar.js:813-823 uses pako). is_local != 0 mirrors the `isLocal` argument used by applyLocalChange
 * (backend.js:84): the patch then carries actor and seq of the single change. want_patch == 0 is
 * Backend.loadChanges (backend.js:116-121): same state transition, no patch computed; *out is set to NULL. */
int amg_apply_changes(amg_backend* b, const uint8_t* const* bufs, const size_t* lens, size_t n, int is_local, int want_patch,
                      amg_patch** out, amg_error* err);
/* Same, with the n changes stored back to back in one buffer: change i is blob[offsets[i] .. offsets[i+1]) (offsets: host
 * memory). This is the bulk-replay entry point. `blob` may be pinned host memory (uploaded in 16 MB pieces, each piece
 * hashed and decoded while the next one is in flight), pageable host memory (staged through the engine's pinned arena
 * mirror) or DEVICE memory (copied device to device: the bytes are already resident). */
int amg_apply_changes_packed(amg_backend* b, const uint8_t* blob, const uint64_t* offsets, size_t n, int is_local, int want_patch,
                             amg_patch** out, amg_error* err);
/* Backend.getPatch(state) — backend/backend.js:127-129 -> new.js:2060-2068 / documentPatch new.js:1604-1635 */
int amg_get_patch(amg_backend* b, amg_patch** out, amg_error* err);

/* header-only patch: maxOp, clock, deps (heads), pendingChanges and the actor table, without diffs */
int amg_get_state(amg_backend* b, amg_patch** out, amg_error* err);

/* Backend.getHeads — backend.js:135-137: n hashes of 32 bytes, ascending */
/* Backend.save (backend/backend.js:93-95, new.js:2033-2055): one buffer = the document chunk */
int amg_save(amg_backend* b, amg_buffers** out, amg_error* err);
int amg_get_heads(amg_backend* b, amg_buffers** out, amg_error* err);
/* Backend.getAllChanges / getChanges(haveDeps) — backend.js:142-156 -> new.js:1921-1973; have_deps = n hashes x 32 bytes */
int amg_get_changes(amg_backend* b, const uint8_t* have_deps, size_t n, amg_buffers** out, amg_error* err);
/* Backend.getChangesAdded(old, new) — backend.js:166-168 -> new.js:1979-1997 */
int amg_get_changes_added(amg_backend* b_new, amg_backend* b_old, amg_buffers** out, amg_error* err);
/* Backend.getChangeByHash — backend.js:176-178 -> new.js:1999-2002; zero buffers when unknown */
int amg_get_change_by_hash(amg_backend* b, const uint8_t hash[32], amg_buffers** out, amg_error* err);
/* Backend.getMissingDeps — backend.js:190-192 -> new.js:2014-2028: hashes of 32 bytes */
int amg_get_missing_deps(amg_backend* b, const uint8_t* heads, size_t n, amg_buffers** out, amg_error* err);
/* state needed by the host-side applyLocalChange (backend.js:54-91): clock[actor] and hashesByActor[actor][index] */
int amg_clock_of(amg_backend* b, const uint8_t* actor, size_t actor_len, uint64_t* seq_out, amg_error* err);
int amg_hash_by_actor(amg_backend* b, const uint8_t* actor, size_t actor_len, uint64_t index, uint8_t hash_out[32], int* found, amg_error* err);

/* returned buffer lists */
size_t amg_buffers_count(const amg_buffers* l);
const uint8_t* amg_buffers_get(const amg_buffers* l, size_t i, size_t* len);
void amg_buffers_free(amg_buffers* l);

/* Patch layout (little endian). amg_patch_bytes returns a header of 20 uint64:
 *   [0] 0x31504747414d41 ("AMAGGP1") [1] maxOp [2] pendingChanges [3] hasActorSeq [4] seq [5] actorOff [6] actorLen
 *   [7] actorsOff [8] nActors [9] clockOff [10] nClock [11] depsOff [12] nDeps [13] propsOff [14] nProps
 *   [15] editsOff [16] nEdits [17] editElemOff [18] bytesOff [19] bytesLen
 * followed by the sections (offsets relative to the start of the buffer, 8-byte aligned):
 *   actors : nActors x { uint32 len; bytes[len]; pad to 4 }   (document actor index -> actor id bytes)
 *   clock  : nClock x { uint64 actorIndex; uint64 seq }
 *   deps   : nDeps x 32 bytes
 *   props  : nProps x { uint64 objId; uint64 opId; uint32 keyOff, keyLen, valLen, valOff, flags, pad }   (map entries)
 *   edits  : nEdits x { uint64 objId; uint64 opId; uint32 index, kind, valLen, valOff }                  (list edits, per object in order)
 *   editElem: nEdits x uint64 elemId (offset 0 in the header = section absent: every insert's elemId is its opId)
 *   bytes  : the map keys and value payloads the records refer to, gathered on the device
 * ids are (counter << 16 | actorIndex); objId 0 = _root. keyOff / valOff are offsets INTO THE PATCH BUFFER (its bytes
 * section): a patch is self-contained, no second buffer is needed to read it. valLen is the reference's VALUE_LEN tag
 * (length << 4 | type, columnar.js:46-49); numeric payloads have been validated like the reference's decodeValue does
 * (columnar.js:300-329) - a malformed one makes the call fail with the reference's RangeError.
 * props.flags = action << 8 | 1 if the key has no visible value (reference emits `key: {}`);
 * edits.kind = (0 insert | 1 remove | 2 update) | 0x100 if the edit starts a new run (edits without the bit
 * continue the previous insert as `multi-insert` / add to the previous remove's count, new.js:747-782)
 * | 0x1000 for a counter whose increments were summed: its value is the int64 (valOff << 32 | valLen)
 * | 0x200 if the insert is rendered as `multi-insert` (set on every member of a run, and on a run start whose
 * followers were popped again by appendUpdate, new.js:811-813) | action << 16.
 * The nested Patch object of @types/automerge/index.d.ts:236-316 is assembled from this by the binding. */
/* The bytes live in a pinned buffer owned by the backend: valid until the next call on the same backend. */
const uint8_t* amg_patch_bytes(const amg_patch* p, size_t* len);
void amg_patch_free(amg_patch* p);
/* host copy of the document arena (every change's bytes back to back, inflated copies of DEFLATEd changes behind them);
 * valid until the next mutating call. The engine keeps no host copy of bytes that were handed over in pinned or device
 * memory: the missing part is fetched from the device by this call (and by getChanges & co, which read from it). */
const uint8_t* amg_arena(amg_backend* b, size_t* len);

/* ---- parity / measurement hooks (not part of the reference surface) ---- */
/* decoded rows of a batch of changes (SoA dump of the decode kernels' output, gathered into batch order), for parity tests
 * (SURVEY.md 8c "parity definition" items 1-2: per-change hash, decoded rows). rows_out (malloc'ed, amg_free_mem) holds 12
 * columns x total_ops {objActor,objCtr,keyActor,keyCtr,keyStrOff,keyStrLen,insert,action,valLen,valOff,predNum,predOff}
 * (change-local actor indexes, 0xffffffff = null, offsets relative to the staged copy of the batch) followed by
 * 2 columns x total_preds {predActor, predCtr}. The document is not touched. */
int amg_debug_decode(amg_backend* b, const uint8_t* blob, const uint64_t* offsets, size_t n, uint8_t* hashes_out /* n*32 */,
                     uint32_t* n_ops_out /* n */, uint32_t** rows_out, size_t* total_ops, size_t* total_preds, amg_error* err);
/* document-ordered op table: rows[n][8] = {objCtr,objActor,idCtr,idActor,keyCtr,keyActor,flags,succNum}; succ[m][2] = {ctr, actor} */
int amg_debug_dump_ops(amg_backend* b, uint64_t** rows_out, size_t* n, uint64_t** succ_out, size_t* m, amg_error* err);
/* timings of the last applyChanges call: [0..11] CUDA-event phases in ms on the engine's main stream (upload + hash + decode of
 * the pieces, inflate + rest of the decode, gate, actors + seq + row finalisation, op set, patch groups + props, list index,
 * edits + copy-out, heads + commit), [12..23] host wall-clock marks (ms since the call started; [23] = the whole ABI call) */
int amg_last_timings(amg_backend* b, float* ms_out, int n);
uint64_t amg_kernel_launches(amg_backend* b);
/* labelled host wall-clock marks of the last applyChanges call ("label=ms ..."), development aid */
size_t amg_debug_marks(amg_backend* b, char* buf, size_t cap);
void amg_free_mem(void* p);
/* one column of a document through either decoder of the load path: kind 0 = RLE of unsigned numbers, 1 = RLE of signed
 * numbers, 2 = delta column, 3 = boolean column; out[n] (INT64_MIN = null). parallel = 1: the token / record decoder for
 * long columns (returns 1 without touching out when it declines the stream); parallel = 0: the serial walker (malformed
 * input is an error, as in the reference). For differential tests of the two. */
int amg_debug_decode_column(amg_backend* b, const uint8_t* bytes, size_t len, int kind, size_t n, int parallel, int64_t* out, amg_error* err);
/* device-only re-run of the decode kernels over the last batch (inputs resident in HBM), for the roofline measurement:
 * ms_sha = SHA-256 kernel, ms_parse = fused decode (k_decode_tiles + k_decode_direct), ms_decode = DecodeColumnKernel over
 * changes of more than 16 ops (0 if none); algo_bytes = SURVEY.md 8d: encoded bytes + 48 B/op + 8 B/pred + 96 B/change */
int amg_bench_decode(amg_backend* b, int iters, float* ms_sha, float* ms_parse, float* ms_decode, uint64_t* algo_bytes, amg_error* err);

#ifdef __cplusplus
}
#endif
#endif
