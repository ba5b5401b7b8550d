// cvaegan_b200 - host side of the step-program kernel (mega.cuh): recording, program upload, launch.
//
// train.cu emits every op of a step through launch_mn / launch_dw / emit_* helpers.  While a program is being
// recorded those helpers append an op record here instead of launching a kernel; mk_flush() uploads the program (a
// content-addressed cache of device buffers with pinned host mirrors, so CUDA-graph replays re-read stable memory) and
// launches ONE cooperative kernel for it.
#include "mega.cuh"

namespace cvg {

using mk::OpRec;

bool mk_supported(const Engine& e) {
  auto ok = [](const int* h) { return h[0] <= mk::MAX_C && h[1] <= mk::MAX_C && h[2] <= mk::MAX_C; };
  if (e.Kc != e.K) return false;      // unconditional networks: stand-alone executor only
  return ok(e.eh) && ok(e.gh) && ok(e.dh) && ok(e.ch) && e.F + e.K <= mk::B_MAXN && e.Z <= mk::B_MAXN && 2 * e.Z <= SN_MAXDIM &&
         e.ch[1] <= 8 * mk::MK_LN_F;
}

void mk_set_kernel_attributes() {
  cudaFuncSetAttribute(mk::step_program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mk::SMEM_BYTES);
}

void mk_destroy(Engine& e) {
  for (auto& s : e.mk.slots) {
    if (s.ev) cudaEventDestroy(s.ev);
    if (s.dev) cudaFree(s.dev);
    if (s.host) cudaFreeHost(s.host);
  }
  e.mk.slots.clear();
}

int mk_begin(Engine& e) {
  MkState& m = e.mk;
  if (!m.enabled) return 0;
  if (m.recording) CVG_FAIL("mk_begin: a program is already being recorded");
  m.recording = true;
  m.par_next = false;
  m.ops.clear();
  m.nops = 0;
  m.phase_items = 0;
  m.scratch_off = 0;
  m.n_exchanges = 0;
  m.dcounter = 0;
  for (int i = 0; i < 4; ++i) m.adam_inc[i] = 0;
  m.pending_red.clear();
  m.prep.clear();
  m.prep_off = 0;
  return 0;
}

// Pre-split copy of a GEMM's weight operand: found or created; a K_PREP op is emitted when the copy is stale (first use
// in this program, or the network was updated by Adam since).  The op joins the current phase - it only reads weights,
// which no op of a step writes before its Adam - and the GEMM's own barrier orders it.
int mk_weight_operand(Engine& e, const GemmArgs& g, bool wt, const float** out, bool* emitted) {
  MkState& m = e.mk;
  if (emitted) *emitted = false;
  const int R = g.R < g.a.rows ? g.R : g.a.rows;
  MkPrepSlot* slot = nullptr;
  for (auto& sl : m.prep)
    if (sl.W == g.W && sl.wt == (wt ? 1 : 0) && sl.wcol0 == g.wcol0 && sl.R == R && sl.N == g.N && sl.ldw == g.ldw) { slot = &sl; break; }
  if (!slot) {
    MkPrepSlot sl;
    sl.W = g.W; sl.ldw = g.ldw; sl.wcol0 = g.wcol0; sl.wt = wt ? 1 : 0; sl.R = R; sl.N = g.N;
    for (int net = 0; net < 4; ++net)
      if (g.W >= e.buf[net].params && g.W < e.buf[net].params + e.lay[net].n_param) sl.net = net;
    if (sl.net < 0) CVG_FAIL("step program: GEMM weight operand outside the bound parameter buffers");
    const long long chunks = (long long)((g.N + 127) / 128) * ((R + mk::KC - 1) / mk::KC);
    sl.off = m.prep_off;
    m.prep_off += chunks * mk::CHUNK_FLOATS;
    if (m.prep_off > e.ws.mk_wprep_floats) CVG_FAIL("step program: pre-split weight buffer exhausted");
    m.prep.push_back(sl);
    slot = &m.prep.back();
  }
  if (!slot->fresh) {
    // append to the newest prep op of the current phase when there is room, else start a new op in this phase
    const int chunks = ((slot->N + 127) / 128) * ((slot->R + mk::KC - 1) / mk::KC);
    mk::PrepEntry pe;
    memset(&pe, 0, sizeof(pe));
    pe.W = slot->W; pe.ldw = slot->ldw; pe.wcol0 = slot->wcol0; pe.wt = slot->wt; pe.R = slot->R; pe.N = slot->N; pe.off = slot->off;
    OpRec* last = m.nops > 0 ? reinterpret_cast<OpRec*>(m.ops.data() + (size_t)(m.nops - 1) * sizeof(OpRec)) : nullptr;
    bool appended = false;
    if (last && last->kind == mk::K_PREP && m.prep_open) {
      mk::PrepArgs a;
      memcpy(&a, last->payload, sizeof(a));
      if (a.n < mk::PREP_MAX) {
        a.e[a.n++] = pe;
        memcpy(last->payload, &a, sizeof(a));
        last->items += chunks;
        m.phase_items += chunks;
        appended = true;
      }
    }
    if (!appended) {
      mk::PrepArgs a;
      memset(&a, 0, sizeof(a));
      a.n = 1;
      a.e[0] = pe;
      a.wprep = e.ws.mk_wprep;
      if (m.nops > 0) m.par_next = true;
      CVG_TRY(mk_push(e, mk::K_PREP, &a, sizeof(a), chunks));
      m.prep_open = true;
    }
    slot->fresh = true;
    if (emitted) *emitted = true;
  }
  *out = e.ws.mk_wprep + slot->off;
  return 0;
}

void mk_weights_updated(Engine& e, int net_mask) {
  for (auto& sl : e.mk.prep)
    if (net_mask & (1 << sl.net)) sl.fresh = false;
}

int mk_push(Engine& e, int kind, const void* payload, size_t bytes, int items, int a0, int a1, int a2, int a3, const void* extra,
            size_t extra_bytes) {
  MkState& m = e.mk;
  if (bytes + extra_bytes > sizeof(OpRec::payload)) CVG_FAIL("mk_push: payload too large");
  const bool par = m.par_next && !m.allbar && m.nops > 0;
  m.par_next = false;
  if (kind != mk::K_PREP) m.prep_open = false;
  if (m.max_ops >= 0 && m.nops >= m.max_ops) return 0;
  if (items <= 0) return 0;
  OpRec r;
  memset(&r, 0, sizeof(r));
  r.kind = kind;
  r.bar_before = (m.nops == 0) ? 0 : (par ? 0 : 1);
  if (r.bar_before || m.nops == 0) m.phase_items = 0;
  r.items = items;
  r.first = m.phase_items;
  m.phase_items += items;
  r.aux[0] = a0; r.aux[1] = a1; r.aux[2] = a2; r.aux[3] = a3;
  if (payload) memcpy(r.payload, payload, bytes);
  if (extra) memcpy(r.payload + bytes, extra, extra_bytes);
  const size_t off = m.ops.size();
  m.ops.resize(off + sizeof(OpRec));
  memcpy(m.ops.data() + off, &r, sizeof(OpRec));
  m.nops++;
  return 0;
}

// Row slices: 256 batch rows per work item (8 staged chunks), at most 32 slices and never more than the scratch holds.
int mk_push_dw(Engine& e, const DwArgs& g0) {
  MkState& m = e.mk;
  DwArgs g = g0;
  if (g.K > mk::B_MAXN) CVG_FAIL("step program: weight-gradient operand wider than 256 features");
  const int kp = (g.K + 15) & ~15;
  int nsplit = (g.M + 255) / 256;
  if (nsplit > 32) nsplit = 32;
  if (nsplit < 1) nsplit = 1;
  for (;;) {
    const long long need = ((long long)g.npass * nsplit * g.N * (kp + 1) + 3) & ~3ll;
    if (m.scratch_off + need <= e.ws.dw_scratch_floats || nsplit == 1) break;
    nsplit = (nsplit + 1) / 2;
  }
  int rows = (g.M + nsplit - 1) / nsplit;
  rows = ((rows + mk::KC - 1) / mk::KC) * mk::KC;
  nsplit = (g.M + rows - 1) / rows;
  g.rows_per_cta = rows;
  const int nz = g.npass * nsplit;
  const long long need = ((long long)nz * g.N * (kp + 1) + 3) & ~3ll;     // keeps every slot 16-byte aligned
  if (m.scratch_off + need > e.ws.dw_scratch_floats) CVG_FAIL("step program: weight-gradient scratch exhausted");
  float* part = e.ws.dw_scratch + m.scratch_off;
  float* bpart = part + (long long)nz * g.N * kp;
  m.scratch_off += need;
  float* ptrs[2] = {part, bpart};
  const int ntile = (g.N + 127) / 128;
  CVG_TRY(mk_push(e, mk::K_DW, &g, sizeof(g), ntile * nz, nsplit, kp, 0, 0, ptrs, sizeof(ptrs)));
  mk::DwRedArgs r;
  memset(&r, 0, sizeof(r));
  r.part = part; r.bpart = bpart;
  r.nz = nz; r.nsplit = nsplit; r.npass = g.npass;
  r.N = g.N; r.K = g.K; r.Kp = kp;
  r.dW = g.dW; r.sdW = g.sdW; r.ldw = g.ldw; r.wcol0 = g.wcol0;
  r.db = g.db; r.label_col = g.label_col;
  r.dgamma = g.dgamma; r.dbeta = g.dbeta;
  r.bstats = g.p.bn.bstats; r.sb = g.p.bn.sb; r.C = g.p.bn.C; r.add_affine = g.add_affine;
  MkPendingRed pr;
  static_assert(sizeof(mk::DwRedArgs) <= sizeof(pr.bytes), "DwRedArgs size");
  memcpy(pr.bytes, &r, sizeof(r));
  pr.items = (int)(((long long)g.N * g.K + 4 * mk::THREADS - 1) / (4 * mk::THREADS));
  if (pr.items < 1) pr.items = 1;
  m.pending_red.push_back(pr);
  return 0;
}

int mk_emit_dwred(Engine& e) {
  MkState& m = e.mk;
  bool first = true;
  for (auto& pr : m.pending_red) {
    if (!first) m.par_next = true;
    CVG_TRY(mk_push(e, mk::K_DWRED, pr.bytes, sizeof(mk::DwRedArgs), pr.items));
    first = false;
  }
  m.pending_red.clear();
  return 0;
}

static unsigned long long fnv1a(const unsigned char* p, size_t n) {
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

constexpr int MK_SLOTS = 32;
constexpr size_t MK_SLOT_BYTES = 384 * 1024;

// Program buffers (device + pinned host mirror): allocated when the workspace is bound, i.e. outside any stream capture
int mk_alloc_slots(Engine& e) {
  MkState& m = e.mk;
  if (!m.slots.empty()) return 0;
  m.slots.resize(MK_SLOTS);
  for (auto& s : m.slots) {
    CVG_CUDA(cudaMalloc(&s.dev, MK_SLOT_BYTES));          // program text, not tensor memory
    CVG_CUDA(cudaMallocHost(&s.host, MK_SLOT_BYTES));
    CVG_CUDA(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  }
  return 0;
}

static int mk_program_slot(Engine& e, cudaStream_t st, const void** dev_out) {
  MkState& m = e.mk;
  const size_t bytes = m.ops.size();
  if (bytes > MK_SLOT_BYTES) CVG_FAIL("step program too long");
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CVG_CUDA(cudaStreamIsCapturing(st, &cap));
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  if (m.slots.empty()) {
    if (capturing) CVG_FAIL("step program: program buffers are not allocated (cvg_bind_workspace does it) and a stream capture is active");
    CVG_TRY(mk_alloc_slots(e));
  }
  const unsigned long long h = fnv1a(m.ops.data(), bytes);
  ++m.clock;
  for (auto& s : m.slots) {
    if (s.bytes == bytes && s.hash == h && memcmp(s.host, m.ops.data(), bytes) == 0) {
      s.last_use = m.clock;
      if (capturing) s.in_graph = true;
      *dev_out = s.dev;
      return 0;
    }
  }
  MkSlot* victim = nullptr;
  for (auto& s : m.slots) {
    if (s.in_graph) continue;
    if (!victim || s.last_use < victim->last_use) victim = &s;
  }
  if (!victim) CVG_FAIL("step program: all program buffers are owned by captured graphs");
  if (victim->bytes != 0) {
    if (capturing) {
      if (cudaEventQuery(victim->ev) != cudaSuccess) CVG_FAIL("step program: program buffer still in flight during stream capture");
    } else {
      CVG_CUDA(cudaEventSynchronize(victim->ev));           // its last upload may still be reading the pinned mirror
    }
  }
  memcpy(victim->host, m.ops.data(), bytes);
  victim->bytes = bytes;
  victim->hash = h;
  victim->last_use = m.clock;
  victim->in_graph = capturing;
  CVG_CUDA(cudaMemcpyAsync(victim->dev, victim->host, bytes, cudaMemcpyHostToDevice, st));
  if (!capturing) CVG_CUDA(cudaEventRecord(victim->ev, st));
  *dev_out = victim->dev;
  return 0;
}

int mk_flush(Engine& e, cudaStream_t st) {
  MkState& m = e.mk;
  if (!m.recording) return 0;
  m.recording = false;
  if (!m.pending_red.empty()) CVG_FAIL("step program: pending weight-gradient reductions at flush");
  if (m.nops == 0) return 0;
  mk::FinishArgs f;
  memset(&f, 0, sizeof(f));
  f.ctl = e.ws.ctl;
  f.dcounter = m.dcounter;
  for (int i = 0; i < 4; ++i) f.adam_inc[i] = m.adam_inc[i];
  f.n_exchanges = m.n_exchanges;
  {
    const int keep = m.max_ops;
    m.max_ops = -1;                 // the finish op is never truncated away
    m.recording = true;
    const int r = mk_push(e, mk::K_FINISH, &f, sizeof(f), 1);
    m.recording = false;
    m.max_ops = keep;
    if (r) return r;
  }
  const void* dev = nullptr;
  CVG_TRY(mk_program_slot(e, st, &dev));
  CVG_CUDA(cudaMemsetAsync(e.ws.mk_bar, 0, sizeof(unsigned int), st));
  mk::Params P;
  memset(&P, 0, sizeof(P));
  P.ops = reinterpret_cast<const OpRec*>(dev);
  P.nops = m.nops;
  P.bar_counter = e.ws.mk_bar;
  if (e.world > 1 && e.nvl.on) P.nvl = e.nvl.dev;
  P.dbg = (getenv("CVG_MK_DBG") && m.nops <= 1024) ? e.ws.mk_dbg : nullptr;     // [0, 1024): cycles per op, [1024, 2048): start clocks
  m.last_kinds.clear();
  for (int i = 0; i < m.nops; ++i) {
    const OpRec* r = reinterpret_cast<const OpRec*>(m.ops.data() + (size_t)i * sizeof(OpRec));
    m.last_kinds.push_back(r->kind | (r->bar_before << 8) | (r->items << 16));
  }
  P.prof = (P.dbg && getenv("CVG_MK_PROF")) ? e.ws.mk_dbg + 2048 : nullptr;
  if (P.prof) CVG_CUDA(cudaMemsetAsync(P.prof, 0, 64 * sizeof(long long), st));
  m.last_nops = m.nops;
  void* args[1] = {&P};
  cudaError_t err;
  if (m.coop) {
    err = cudaLaunchCooperativeKernel((const void*)mk::step_program_kernel, dim3(e.num_sms), dim3(mk::THREADS), args, mk::SMEM_BYTES, st);
  } else {
    mk::step_program_kernel<<<e.num_sms, mk::THREADS, mk::SMEM_BYTES, st>>>(P);
    err = cudaGetLastError();
  }
  if (err != cudaSuccess) CVG_FAIL(std::string("step program launch: ") + cudaGetErrorString(err));
  e.launches++;
  return 0;
}

}  // namespace cvg
