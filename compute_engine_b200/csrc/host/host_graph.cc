#include "host_graph.h"

#include "builtin_params.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace lce_b200 {

namespace {
size_t TypeSize(TfLiteType t) {
  switch (t) {
    case kTfLiteFloat32: return 4;
    case kTfLiteInt32: return 4;
    case kTfLiteInt64: return 8;
    case kTfLiteUInt8:
    case kTfLiteInt8:
    case kTfLiteBool: return 1;
    default: return 0;
  }
}
TfLiteIntArray* MakeDims(const std::vector<int>& d) {
  TfLiteIntArray* a = LceB200IntArrayCreate(static_cast<int>(d.size()));
  for (size_t i = 0; i < d.size(); ++i) a->data[i] = d[i];
  return a;
}
size_t BytesOf(const TfLiteTensor& t) {
  size_t n = TypeSize(t.type);
  for (int i = 0; i < t.dims->size; ++i) n *= static_cast<size_t>(std::max(t.dims->data[i], 0));
  return n;
}
constexpr size_t kAlign = 256;
size_t AlignUp(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }
}  // namespace

// ------------------------------ OpResolver ------------------------------- //
void OpResolver::AddCustom(const char* name, const TfLiteRegistration* r) { custom_[name] = r; }
void OpResolver::AddBuiltin(int code, const TfLiteRegistration* r) { builtin_[code] = r; }
const TfLiteRegistration* OpResolver::FindCustom(const std::string& name) const {
  auto it = custom_.find(name);
  return it == custom_.end() ? nullptr : it->second;
}
const TfLiteRegistration* OpResolver::FindBuiltin(int code) const {
  auto it = builtin_.find(code);
  return it == builtin_.end() ? nullptr : it->second;
}

// --------------------------------- Graph --------------------------------- //
Graph::Graph(bool device_arena) : device_arena_(device_arena) {
  memset(&ctx_, 0, sizeof(ctx_));
  ctx_.impl_ = this;
  ctx_.ResizeTensor = &Graph::ResizeTensorCb;
  ctx_.ReportError = &Graph::ReportErrorCb;
  ctx_.AddTensors = &Graph::AddTensorsCb;
  ctx_.GetTensor = &Graph::GetTensorCb;
  ctx_.GetExternalContext = &Graph::GetExternalContextCb;
  ctx_.SetExternalContext = &Graph::SetExternalContextCb;
  ctx_.recommended_num_threads = 1;
  if (device_arena_) {
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess) stream_ = s;
    else cudaGetLastError();
  }
}

Graph::~Graph() {
  for (auto& n : nodes_) {
    if (n->initialized && n->registration->free) n->registration->free(&ctx_, n->node.user_data);
    LceB200IntArrayFree(n->node.inputs);
    LceB200IntArrayFree(n->node.outputs);
    LceB200IntArrayFree(n->node.temporaries);
    LceB200IntArrayFree(n->node.intermediates);
  }
  if (graph_exec_) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec_));
  FreeArena();
  for (void* p : const_dev_)
    if (p) cudaFree(p);
  for (auto& t : tensors_) {
    LceB200IntArrayFree(t.dims);
    if (t.quantization.params) {
      auto* q = static_cast<TfLiteAffineQuantization*>(t.quantization.params);
      free(q->scale);
      LceB200IntArrayFree(q->zero_point);
      free(q);
    }
  }
  if (stream_) cudaStreamDestroy(static_cast<cudaStream_t>(stream_));
}

void Graph::RefreshContext() {
  ctx_.tensors = tensors_.data();
  ctx_.tensors_size = tensors_.size();
}

int Graph::AddTensor(TfLiteType type, const std::vector<int>& dims, const void* const_data,
                     size_t const_bytes, bool has_quant, float scale, int zero_point,
                     const std::string& name, bool const_on_device) {
  TfLiteTensor t;
  memset(&t, 0, sizeof(t));
  t.type = type;
  t.dims = MakeDims(dims);
  t.params.scale = scale;
  t.params.zero_point = zero_point;
  if (has_quant) {
    auto* q = static_cast<TfLiteAffineQuantization*>(malloc(sizeof(TfLiteAffineQuantization)));
    q->scale = static_cast<TfLiteFloatArray*>(malloc(sizeof(TfLiteFloatArray) + sizeof(float)));
    q->scale->size = 1;
    q->scale->data[0] = scale;
    q->zero_point = LceB200IntArrayCreate(1);
    q->zero_point->data[0] = zero_point;
    q->quantized_dimension = 0;
    t.quantization.type = kTfLiteAffineQuantization;
    t.quantization.params = q;
  }
  t.bytes = BytesOf(t);
  names_.push_back(name);
  const_host_.emplace_back();
  const_dev_.push_back(nullptr);
  const int idx = static_cast<int>(tensors_.size());
  if (const_data) {
    auto& h = const_host_.back();
    h.assign(static_cast<const uint8_t*>(const_data),
             static_cast<const uint8_t*>(const_data) + const_bytes);
    h.resize(const_bytes + 64, 0);  // slack: kernels may read whole vectors
    t.allocation_type = kTfLiteMmapRo;
    t.data.raw = reinterpret_cast<char*>(h.data());
    t.bytes = const_bytes;
    if (const_on_device && device_arena_) {
      void* d = nullptr;
      if (cudaMalloc(&d, const_bytes + 64) == cudaSuccess &&
          cudaMemcpy(d, h.data(), const_bytes + 64, cudaMemcpyHostToDevice) == cudaSuccess) {
        const_dev_.back() = d;
        t.data.raw = static_cast<char*>(d);
      } else {
        cudaGetLastError();
      }
    }
  } else {
    t.allocation_type = kTfLiteArenaRw;
  }
  tensors_.push_back(t);
  tensors_.back().name = names_.back().c_str();
  for (size_t i = 0; i < tensors_.size(); ++i) tensors_[i].name = names_[i].c_str();
  RefreshContext();
  allocated_ = false;
  return idx;
}

int Graph::AddNode(const TfLiteRegistration* registration, const std::vector<int>& inputs,
                   const std::vector<int>& outputs, const uint8_t* custom_options,
                   size_t options_len, const void* builtin_data, size_t builtin_bytes,
                   const std::string& name) {
  auto rec = std::make_unique<NodeRecord>();
  rec->registration = registration;
  rec->name = name;
  if (custom_options) rec->custom_options.assign(custom_options, custom_options + options_len);
  if (builtin_data)
    rec->builtin_blob.assign(static_cast<const uint8_t*>(builtin_data),
                             static_cast<const uint8_t*>(builtin_data) + builtin_bytes);
  rec->node.inputs = MakeDims(inputs);
  rec->node.outputs = MakeDims(outputs);
  rec->node.temporaries = LceB200IntArrayCreate(0);
  rec->node.intermediates = LceB200IntArrayCreate(0);
  rec->node.custom_initial_data = rec->custom_options.data();
  rec->node.custom_initial_data_size = static_cast<int>(rec->custom_options.size());
  rec->node.builtin_data = rec->builtin_blob.empty() ? nullptr : rec->builtin_blob.data();
  nodes_.push_back(std::move(rec));
  allocated_ = false;
  return static_cast<int>(nodes_.size()) - 1;
}

TfLiteStatus Graph::ResizeTensorCb(TfLiteContext* c, TfLiteTensor* t, TfLiteIntArray* new_size) {
  // Takes ownership of new_size (common.h: ResizeTensor contract).
  auto* g = static_cast<Graph*>(c->impl_);
  if (t->allocation_type == kTfLiteMmapRo) {
    LceB200IntArrayFree(new_size);
    g->error_ = "ResizeTensor called on a constant tensor";
    return kTfLiteError;
  }
  LceB200IntArrayFree(t->dims);
  t->dims = new_size;
  t->bytes = BytesOf(*t);
  g->allocated_ = false;
  return kTfLiteOk;
}

void Graph::ReportErrorCb(TfLiteContext* c, const char* msg, ...) {
  auto* g = static_cast<Graph*>(c->impl_);
  char buf[1024];
  va_list ap;
  va_start(ap, msg);
  vsnprintf(buf, sizeof(buf), msg, ap);
  va_end(ap);
  g->error_ = buf;
}

TfLiteStatus Graph::AddTensorsCb(TfLiteContext* c, int n, int* first) {
  auto* g = static_cast<Graph*>(c->impl_);
  if (first) *first = static_cast<int>(g->tensors_.size());
  for (int i = 0; i < n; ++i)
    g->AddTensor(kTfLiteNoType, {0}, nullptr, 0, false, 0.f, 0, "temporary");
  return kTfLiteOk;
}

TfLiteTensor* Graph::GetTensorCb(const TfLiteContext* c, int i) {
  auto* g = static_cast<Graph*>(c->impl_);
  if (i < 0 || i >= static_cast<int>(g->tensors_.size())) return nullptr;
  return &g->tensors_[i];
}

TfLiteExternalContext* Graph::GetExternalContextCb(TfLiteContext* c, TfLiteExternalContextType t) {
  auto* g = static_cast<Graph*>(c->impl_);
  return (t >= 0 && t < kTfLiteMaxExternalContexts) ? g->external_[t] : nullptr;
}
void Graph::SetExternalContextCb(TfLiteContext* c, TfLiteExternalContextType t,
                                 TfLiteExternalContext* e) {
  auto* g = static_cast<Graph*>(c->impl_);
  if (t >= 0 && t < kTfLiteMaxExternalContexts) g->external_[t] = e;
}

void Graph::FreeArena() {
  if (!arena_) return;
  if (device_arena_) cudaFree(arena_);
  else free(arena_);
  arena_ = nullptr;
  arena_bytes_ = 0;
}

// Greedy-by-size offset assignment over tensor lifetimes (the idea of TFLite's
// arena planner): tensors whose [first_use, last_use] node ranges are disjoint may
// share bytes. Graph inputs live from -1, graph outputs until the end.
TfLiteStatus Graph::PlanArena() {
  const int n_nodes = static_cast<int>(nodes_.size());
  const int n_t = static_cast<int>(tensors_.size());
  std::vector<int> first(n_t, n_nodes + 1), last(n_t, -2);
  auto touch = [&](int t, int when) {
    if (t < 0 || t >= n_t) return;
    first[t] = std::min(first[t], when);
    last[t] = std::max(last[t], when);
  };
  for (int t : inputs_) touch(t, -1);
  for (int i = 0; i < n_nodes; ++i) {
    const TfLiteNode& nd = nodes_[i]->node;
    for (int k = 0; k < nd.inputs->size; ++k) touch(nd.inputs->data[k], i);
    for (int k = 0; k < nd.outputs->size; ++k) touch(nd.outputs->data[k], i);
    for (int k = 0; k < nd.temporaries->size; ++k) touch(nd.temporaries->data[k], i);
  }
  for (int t : outputs_) touch(t, n_nodes);
  if (preserve_all_)
    for (int t = 0; t < n_t; ++t)
      if (last[t] >= -1) { first[t] = -1; last[t] = n_nodes; }
  struct Item { int t; size_t bytes; size_t off; };
  std::vector<Item> items;
  for (int t = 0; t < n_t; ++t) {
    if (tensors_[t].allocation_type != kTfLiteArenaRw || last[t] < -1) continue;
    items.push_back({t, AlignUp(std::max<size_t>(tensors_[t].bytes, 1) + 64), 0});
  }
  std::sort(items.begin(), items.end(),
            [](const Item& a, const Item& b) { return a.bytes > b.bytes; });
  std::vector<Item> placed;
  size_t total = 0;
  for (auto& it : items) {
    std::vector<std::pair<size_t, size_t>> busy;
    for (auto& p : placed)
      if (!(last[p.t] < first[it.t] || last[it.t] < first[p.t]))
        busy.emplace_back(p.off, p.off + p.bytes);
    std::sort(busy.begin(), busy.end());
    size_t off = 0;
    for (auto& b : busy) {
      if (off + it.bytes <= b.first) break;
      off = std::max(off, b.second);
    }
    it.off = off;
    placed.push_back(it);
    total = std::max(total, off + it.bytes);
  }
  if (total > arena_bytes_) {
    FreeArena();
    if (device_arena_) {
      if (cudaMalloc(&arena_, total) != cudaSuccess) {
        error_ = std::string("cudaMalloc of the tensor arena failed: ") +
                 cudaGetErrorString(cudaGetLastError());
        return kTfLiteError;
      }
    } else {
      arena_ = malloc(total);
      if (!arena_) return kTfLiteError;
    }
    arena_bytes_ = total;
  }
  for (auto& p : placed) tensors_[p.t].data.raw = static_cast<char*>(arena_) + p.off;
  return kTfLiteOk;
}

extern "C" TfLiteRegistration* lce_b200_internal_Register_BCONV_2D_FUSED(const TfLiteRegistration* original);

int Graph::FuseResidualBlocks() {
  if (!device_arena_) return 0;
  int removed = 0;
  auto consumers = [&](int t) {
    std::vector<size_t> c;
    for (size_t i = 0; i < nodes_.size(); ++i) {
      const TfLiteIntArray* in = nodes_[i]->node.inputs;
      for (int k = 0; k < in->size; ++k)
        if (in->data[k] == t) { c.push_back(i); break; }
    }
    return c;
  };
  auto is_graph_output = [&](int t) {
    return std::find(outputs_.begin(), outputs_.end(), t) != outputs_.end();
  };
  for (size_t i = 0; i < nodes_.size(); ++i) {
    NodeRecord& b = *nodes_[i];
    if (b.name != "LceBconv2d" || b.initialized || b.node.inputs->size != 5) continue;
    TfLiteRegistration* fused_reg = lce_b200_internal_Register_BCONV_2D_FUSED(b.registration);
    if (fused_reg == nullptr) continue;   // a foreign "LceBconv2d" (e.g. the CPU test double)
    const int y = b.node.outputs->data[0];
    if (tensors_[y].type != kTfLiteFloat32 || is_graph_output(y)) continue;
    std::vector<size_t> cy = consumers(y);
    if (cy.size() != 1) continue;
    const size_t ai = cy[0];
    NodeRecord& a = *nodes_[ai];
    if (a.name != "builtin:0" || a.node.inputs->size != 2 || ai <= i) continue;  // ADD
    const int in0 = a.node.inputs->data[0], in1 = a.node.inputs->data[1];
    if (in0 == in1) continue;
    const int r = in0 == y ? in1 : in0;
    if (tensors_[r].type != kTfLiteFloat32 || tensors_[r].allocation_type == kTfLiteMmapRo) continue;
    // Preconditions of the fused kernel, checked before the graph is rewritten (both graphs below
    // run correctly unfused): the shortcut has exactly the convolution's shape (ADD would also
    // accept a broadcast over the last dimension), and the convolution is not grouped (the fused
    // sign-pack needs groups == 1): filter words per pixel == input words per pixel.
    {
      const TfLiteIntArray* ys = tensors_[y].dims;
      const TfLiteIntArray* rs = tensors_[r].dims;
      bool same = ys && rs && ys->size == rs->size;
      for (int k = 0; same && k < ys->size; ++k) same = ys->data[k] == rs->data[k];
      if (!same) continue;
      const TfLiteIntArray* xs = tensors_[b.node.inputs->data[0]].dims;
      const TfLiteIntArray* fs = tensors_[b.node.inputs->data[1]].dims;
      if (!xs || !fs || xs->size != 4 || fs->size != 4 || xs->data[3] != fs->data[3]) continue;
    }
    // the shortcut must already exist when the bconv runs
    bool r_ready = std::find(inputs_.begin(), inputs_.end(), r) != inputs_.end();
    for (size_t k = 0; k < i && !r_ready; ++k)
      for (int q = 0; q < nodes_[k]->node.outputs->size; ++q)
        if (nodes_[k]->node.outputs->data[q] == r) r_ready = true;
    if (!r_ready) continue;
    const int o = a.node.outputs->data[0];
    int32_t add_act = 0;
    if (a.builtin_blob.size() >= sizeof(BuiltinParams))
      add_act = reinterpret_cast<const BuiltinParams*>(a.builtin_blob.data())->activation;
    // optional: the LceQuantize that consumes the sum
    int packed = -1;
    size_t qi = nodes_.size();
    for (size_t c : consumers(o)) {
      NodeRecord& q = *nodes_[c];
      if (q.name == "LceQuantize" && c > ai && tensors_[o].type == kTfLiteFloat32) {
        packed = q.node.outputs->data[0];
        qi = c;
        break;
      }
    }
    // rewrite node i
    std::vector<int> ins(b.node.inputs->data, b.node.inputs->data + 5);
    ins.push_back(r);
    std::vector<int> outs{o};
    if (packed >= 0) outs.push_back(packed);
    LceB200IntArrayFree(b.node.inputs);
    LceB200IntArrayFree(b.node.outputs);
    b.node.inputs = MakeDims(ins);
    b.node.outputs = MakeDims(outs);
    b.registration = fused_reg;
    b.builtin_blob.assign(reinterpret_cast<const uint8_t*>(&add_act),
                          reinterpret_cast<const uint8_t*>(&add_act) + 4);
    b.node.builtin_data = b.builtin_blob.data();
    b.name = packed >= 0 ? "LceBconv2d+ADD+LceQuantize" : "LceBconv2d+ADD";
    // drop the absorbed nodes (higher index first)
    auto drop = [&](size_t idx) {
      NodeRecord& n = *nodes_[idx];
      LceB200IntArrayFree(n.node.inputs);
      LceB200IntArrayFree(n.node.outputs);
      LceB200IntArrayFree(n.node.temporaries);
      LceB200IntArrayFree(n.node.intermediates);
      nodes_.erase(nodes_.begin() + idx);
      ++removed;
    };
    if (qi < nodes_.size()) drop(qi);
    drop(ai);
  }
  allocated_ = false;
  return removed;
}

const TfLiteRegistration* FusedPoolDepthwiseRegistration();  // builtin_ops.cc
const TfLiteRegistration* FusedStemRegistration();

// [DEQUANTIZE ->] CONV_2D(3x3, stride 2, 3 -> 16) -> DEPTHWISE_CONV_2D(3x3, stride 2, multiplier 1),
// each intermediate consumed once: one node (QuickNet's stem). The image is read once, as the
// bytes the host shipped; the dequantised image and the first conv's map never reach HBM.
int Graph::FuseStem() {
  auto sole_consumer = [&](int t, size_t after) -> size_t {
    if (std::find(outputs_.begin(), outputs_.end(), t) != outputs_.end()) return nodes_.size();
    size_t found = nodes_.size();
    int n = 0;
    for (size_t k = 0; k < nodes_.size(); ++k)
      for (int q = 0; q < nodes_[k]->node.inputs->size; ++q)
        if (nodes_[k]->node.inputs->data[q] == t) { ++n; found = k; }
    return (n == 1 && found > after && nodes_[found]->node.inputs->data[0] == t) ? found
                                                                                 : nodes_.size();
  };
  auto bp = [](const NodeRecord& r) {
    return reinterpret_cast<const BuiltinParams*>(r.builtin_blob.data());
  };
  auto dims4 = [&](int t, int d) { return tensors_[t].dims->size == 4 ? tensors_[t].dims->data[d] : -1; };
  auto erase_node = [&](size_t victim) {
    NodeRecord& v = *nodes_[victim];
    LceB200IntArrayFree(v.node.inputs);
    LceB200IntArrayFree(v.node.outputs);
    LceB200IntArrayFree(v.node.temporaries);
    LceB200IntArrayFree(v.node.intermediates);
    nodes_.erase(nodes_.begin() + victim);
  };
  for (size_t i = 0; i < nodes_.size(); ++i) {
    NodeRecord& c1 = *nodes_[i];
    if (c1.name != "builtin:3" || c1.initialized || c1.node.inputs->size < 3 ||
        c1.builtin_blob.size() < sizeof(BuiltinParams))
      continue;
    const int w1 = c1.node.inputs->data[1];
    const int x = c1.node.inputs->data[0];
    if (dims4(w1, 0) != 16 || dims4(w1, 1) != 3 || dims4(w1, 2) != 3 || dims4(w1, 3) != 3 ||
        bp(c1)->stride_h != 2 || bp(c1)->stride_w != 2 || bp(c1)->dilation_h != 1 ||
        bp(c1)->dilation_w != 1 || tensors_[x].type != kTfLiteFloat32 ||
        tensors_[c1.node.outputs->data[0]].type != kTfLiteFloat32)
      continue;
    const size_t di = sole_consumer(c1.node.outputs->data[0], i);
    if (di >= nodes_.size()) continue;
    NodeRecord& dw = *nodes_[di];
    if (dw.name != "builtin:4" || dw.initialized || dw.node.inputs->size < 3 ||
        dw.builtin_blob.size() < sizeof(BuiltinParams))
      continue;
    const int w2 = dw.node.inputs->data[1];
    if (dims4(w2, 1) != 3 || dims4(w2, 2) != 3 || dims4(w2, 3) != 16 || bp(dw)->stride_h != 2 ||
        bp(dw)->stride_w != 2 || bp(dw)->dilation_h != 1 || bp(dw)->dilation_w != 1 ||
        bp(dw)->depth_multiplier != 1)
      continue;
    // the kernel takes the filters and biases by value: they must be constants
    bool consts = true;
    for (int t : {w1, c1.node.inputs->data[2], w2, dw.node.inputs->data[2]})
      if (t >= 0 && tensors_[t].allocation_type != kTfLiteMmapRo) consts = false;
    if (!consts) continue;
    // a DEQUANTIZE that feeds only this conv is folded in: the kernel reads the quantised image
    int src = x;
    size_t qi = nodes_.size();
    for (size_t k = 0; k < i; ++k) {
      NodeRecord& dq = *nodes_[k];
      if (dq.name == "builtin:6" && !dq.initialized && dq.node.outputs->size == 1 &&
          dq.node.outputs->data[0] == x && dq.node.inputs->size == 1 && sole_consumer(x, k) == i &&
          (tensors_[dq.node.inputs->data[0]].type == kTfLiteInt8 ||
           tensors_[dq.node.inputs->data[0]].type == kTfLiteUInt8)) {
        qi = k;
        src = dq.node.inputs->data[0];
      }
    }
    std::vector<uint8_t> blob(2 * sizeof(BuiltinParams));
    memcpy(blob.data(), bp(c1), sizeof(BuiltinParams));
    memcpy(blob.data() + sizeof(BuiltinParams), bp(dw), sizeof(BuiltinParams));
    std::vector<int> ins{src, w1, c1.node.inputs->data[2], w2, dw.node.inputs->data[2]};
    std::vector<int> outs{dw.node.outputs->data[0]};
    LceB200IntArrayFree(c1.node.inputs);
    LceB200IntArrayFree(c1.node.outputs);
    c1.node.inputs = MakeDims(ins);
    c1.node.outputs = MakeDims(outs);
    c1.builtin_blob = blob;
    c1.node.builtin_data = c1.builtin_blob.data();
    c1.registration = FusedStemRegistration();
    c1.name = qi < nodes_.size() ? "DEQUANTIZE+CONV_2D+DEPTHWISE_CONV_2D" : "CONV_2D+DEPTHWISE_CONV_2D";
    erase_node(di);                         // the later node first
    int removed = 1;
    if (qi < nodes_.size()) {
      erase_node(qi);
      ++removed;
    }
    allocated_ = false;
    return removed;
  }
  return 0;
}

// CONV_2D (float) whose output also feeds an LceQuantize: the conv emits the packed signs from its
// epilogue as a second output and the LceQuantize node disappears (first layer of every stage).
int Graph::FuseConvQuantize() {
  int removed = 0;
  for (size_t i = 0; i < nodes_.size(); ++i) {
    NodeRecord& conv = *nodes_[i];
    if (conv.name != "builtin:3" || conv.initialized || conv.node.outputs->size != 1) continue;
    const int y = conv.node.outputs->data[0];
    if (tensors_[y].type != kTfLiteFloat32) continue;
    for (size_t k = i + 1; k < nodes_.size(); ++k) {
      NodeRecord& q = *nodes_[k];
      if (q.name != "LceQuantize" || q.initialized || q.node.inputs->size != 1 ||
          q.node.inputs->data[0] != y || q.node.outputs->size != 1)
        continue;
      const int packed = q.node.outputs->data[0];
      if (tensors_[packed].type != kTfLiteInt32) break;
      std::vector<int> outs{y, packed};
      LceB200IntArrayFree(conv.node.outputs);
      conv.node.outputs = MakeDims(outs);
      conv.name = "CONV_2D+LceQuantize";
      LceB200IntArrayFree(q.node.inputs);
      LceB200IntArrayFree(q.node.outputs);
      LceB200IntArrayFree(q.node.temporaries);
      LceB200IntArrayFree(q.node.intermediates);
      nodes_.erase(nodes_.begin() + k);
      ++removed;
      allocated_ = false;
      break;
    }
  }
  return removed;
}

int Graph::FuseFloatGlue() {
  if (!device_arena_) return 0;
  // LCE_B200_FUSE_STEM=0 / LCE_B200_FUSE_CONV_QUANT=0 switch the two fusions off (A/B runs)
  const char* stem_env = getenv("LCE_B200_FUSE_STEM");
  int removed = (stem_env && stem_env[0] == '0') ? 0 : FuseStem();
  // CONV_2D -> LceQuantize: the tensor-core pointwise kernel (lce_b200_pw.cuh) has the output row
  // in registers and emits the sign words for free; other shapes pack with the stand-alone kernel
  const char* cq_env = getenv("LCE_B200_FUSE_CONV_QUANT");
  if (!(cq_env && cq_env[0] == '0')) removed += FuseConvQuantize();
  // RELU -> MEAN (the model heads): the mean applies the activation as it reads
  for (size_t i = 0; i < nodes_.size(); ++i) {
    NodeRecord& relu = *nodes_[i];
    if (relu.name != "builtin:19" || relu.initialized || relu.node.inputs->size != 1 || relu.node.outputs->size != 1)
      continue;
    const int y = relu.node.outputs->data[0];
    if (std::find(outputs_.begin(), outputs_.end(), y) != outputs_.end()) continue;
    size_t mi = nodes_.size();
    int n_cons = 0;
    for (size_t k = 0; k < nodes_.size(); ++k)
      for (int q = 0; q < nodes_[k]->node.inputs->size; ++q)
        if (nodes_[k]->node.inputs->data[q] == y) { ++n_cons; mi = k; }
    if (n_cons != 1 || mi <= i) continue;
    NodeRecord& mean = *nodes_[mi];
    if (mean.name != "builtin:40" || mean.initialized || mean.node.inputs->data[0] != y ||
        mean.builtin_blob.size() < sizeof(BuiltinParams))
      continue;
    reinterpret_cast<BuiltinParams*>(mean.builtin_blob.data())->activation = 1;   // kTfLiteActRelu
    mean.node.builtin_data = mean.builtin_blob.data();
    mean.node.inputs->data[0] = relu.node.inputs->data[0];
    mean.name = "RELU+MEAN";
    LceB200IntArrayFree(relu.node.inputs);
    LceB200IntArrayFree(relu.node.outputs);
    LceB200IntArrayFree(relu.node.temporaries);
    LceB200IntArrayFree(relu.node.intermediates);
    nodes_.erase(nodes_.begin() + i);
    ++removed;
    allocated_ = false;
    break;
  }
  for (size_t i = 0; i < nodes_.size(); ++i) {
    NodeRecord& pool = *nodes_[i];
    if (pool.name != "builtin:17" || pool.initialized ||
        pool.builtin_blob.size() < sizeof(BuiltinParams))
      continue;
    const auto* pp = reinterpret_cast<const BuiltinParams*>(pool.builtin_blob.data());
    if (pp->filter_h != 2 || pp->filter_w != 2 || pp->stride_h != 1 || pp->stride_w != 1 ||
        pp->padding != 1 /* VALID */ || pp->activation != 0)
      continue;
    const int y = pool.node.outputs->data[0];
    if (std::find(outputs_.begin(), outputs_.end(), y) != outputs_.end()) continue;
    size_t di = nodes_.size();
    int n_cons = 0;
    for (size_t k = 0; k < nodes_.size(); ++k)
      for (int q = 0; q < nodes_[k]->node.inputs->size; ++q)
        if (nodes_[k]->node.inputs->data[q] == y) { ++n_cons; di = k; }
    if (n_cons != 1 || di <= i) continue;
    NodeRecord& dw = *nodes_[di];
    if (dw.name != "builtin:4" || dw.node.inputs->data[0] != y ||
        dw.builtin_blob.size() < sizeof(BuiltinParams))
      continue;
    const auto* dp = reinterpret_cast<const BuiltinParams*>(dw.builtin_blob.data());
    const TfLiteTensor& f = tensors_[dw.node.inputs->data[1]];
    if (f.dims->size != 4 || f.dims->data[1] != 3 || f.dims->data[2] != 3 ||
        (f.dims->data[3] & 3) || dp->dilation_h != 1 || dp->dilation_w != 1 ||
        dp->depth_multiplier != 1 || tensors_[y].type != kTfLiteFloat32)
      continue;
    // rewrite the pool node into the fused node; it takes the depthwise node's constants / output
    std::vector<uint8_t> blob(2 * sizeof(BuiltinParams));
    memcpy(blob.data(), pp, sizeof(BuiltinParams));
    memcpy(blob.data() + sizeof(BuiltinParams), dp, sizeof(BuiltinParams));
    std::vector<int> ins{pool.node.inputs->data[0], dw.node.inputs->data[1],
                         dw.node.inputs->size > 2 ? dw.node.inputs->data[2] : -1};
    std::vector<int> outs{dw.node.outputs->data[0]};
    LceB200IntArrayFree(pool.node.inputs);
    LceB200IntArrayFree(pool.node.outputs);
    pool.node.inputs = MakeDims(ins);
    pool.node.outputs = MakeDims(outs);
    pool.builtin_blob = blob;
    pool.node.builtin_data = pool.builtin_blob.data();
    pool.registration = FusedPoolDepthwiseRegistration();
    pool.name = "MAX_POOL_2D+DEPTHWISE_CONV_2D";
    LceB200IntArrayFree(dw.node.inputs);
    LceB200IntArrayFree(dw.node.outputs);
    LceB200IntArrayFree(dw.node.temporaries);
    LceB200IntArrayFree(dw.node.intermediates);
    nodes_.erase(nodes_.begin() + di);
    ++removed;
  }
  allocated_ = false;
  return removed;
}

TfLiteStatus Graph::AllocateTensors() {
  RefreshContext();
  if (graph_exec_) {
    cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec_));
    graph_exec_ = nullptr;
  }
  for (auto& n : nodes_) {
    if (!n->initialized) {
      if (n->registration->init) {
        // custom ops get their flexbuffer options, builtins this host's parameter blob
        const bool custom = !n->custom_options.empty() || !n->node.builtin_data;
        const char* buf = custom ? reinterpret_cast<const char*>(n->custom_options.data())
                                 : static_cast<const char*>(n->node.builtin_data);
        const size_t len = custom ? n->custom_options.size() : n->builtin_blob.size();
        n->node.user_data = n->registration->init(&ctx_, buf, len);
      }
      n->initialized = true;
    }
    if (n->registration->prepare) {
      if (n->registration->prepare(&ctx_, &n->node) != kTfLiteOk) {
        if (error_.empty()) error_ = "prepare failed for node " + n->name;
        return kTfLiteError;
      }
      RefreshContext();
    }
  }
  if (PlanArena() != kTfLiteOk) return kTfLiteError;
  allocated_ = true;
  warmed_ = false;
  return kTfLiteOk;
}

TfLiteStatus Graph::ResizeInputTensor(int t, const std::vector<int>& dims) {
  if (t < 0 || t >= static_cast<int>(tensors_.size())) return kTfLiteError;
  return ResizeTensorCb(&ctx_, &tensors_[t], MakeDims(dims));
}

TfLiteStatus Graph::EnableCudaGraph(bool on) {
  use_cuda_graph_ = on && device_arena_;
  if (!use_cuda_graph_ && graph_exec_) {
    cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec_));
    graph_exec_ = nullptr;
  }
  return kTfLiteOk;
}

TfLiteStatus Graph::Invoke() {
  if (!allocated_ && AllocateTensors() != kTfLiteOk) return kTfLiteError;
  lce_b200_set_stream(stream_);  // nullptr (legacy default stream) for a host arena
  const bool prof = profiling_ && device_arena_ && !use_cuda_graph_;
  if (prof && node_events_.size() != nodes_.size()) {
    node_events_.assign(nodes_.size(), {});
    node_ms_.assign(nodes_.size(), 0.0);
  }
  auto run_nodes = [&]() -> TfLiteStatus {
    for (size_t i = 0; i < nodes_.size(); ++i) {
      auto& n = nodes_[i];
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (prof) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, static_cast<cudaStream_t>(stream_));
      }
      if (n->registration->invoke(&ctx_, &n->node) != kTfLiteOk) {
        if (error_.empty()) error_ = "invoke failed for node " + n->name;
        return kTfLiteError;
      }
      if (prof) {
        cudaEventRecord(e1, static_cast<cudaStream_t>(stream_));
        node_events_[i].emplace_back(e0, e1);
      }
    }
    return kTfLiteOk;
  };
  if (!use_cuda_graph_) return run_nodes();
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!graph_exec_) {
    // First call runs eagerly (ops build their plans: allocations are not
    // capturable); the second call records the launch sequence.
    if (!warmed_) {
      warmed_ = true;
      return run_nodes();
    }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return run_nodes();
    }
    TfLiteStatus rc = run_nodes();
    if (cudaStreamEndCapture(s, &graph) != cudaSuccess || rc != kTfLiteOk || !graph) {
      cudaGetLastError();
      error_ = "CUDA graph capture failed";
      return kTfLiteError;
    }
    cudaGraphExec_t exec = nullptr;
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      cudaGraphDestroy(graph);
      error_ = "cudaGraphInstantiate failed";
      return kTfLiteError;
    }
    cudaGraphDestroy(graph);
    graph_exec_ = exec;
  }
  if (cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec_), s) != cudaSuccess) {
    error_ = "cudaGraphLaunch failed";
    return kTfLiteError;
  }
  return kTfLiteOk;
}

void Graph::ResetProfile() {
  for (auto& v : node_events_)
    for (auto& p : v) {
      cudaEventDestroy(static_cast<cudaEvent_t>(p.first));
      cudaEventDestroy(static_cast<cudaEvent_t>(p.second));
    }
  node_events_.assign(nodes_.size(), {});
  node_ms_.assign(nodes_.size(), 0.0);
}

double Graph::NodeTimeMs(size_t node) {
  if (node >= node_events_.size()) return 0.0;
  if (stream_) cudaStreamSynchronize(static_cast<cudaStream_t>(stream_));
  for (auto& p : node_events_[node]) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(p.first),
                             static_cast<cudaEvent_t>(p.second)) == cudaSuccess)
      node_ms_[node] += ms;
    cudaEventDestroy(static_cast<cudaEvent_t>(p.first));
    cudaEventDestroy(static_cast<cudaEvent_t>(p.second));
  }
  node_events_[node].clear();
  return node_ms_[node];
}

TfLiteStatus Graph::Synchronize() {
  if (device_arena_ && cudaStreamSynchronize(static_cast<cudaStream_t>(stream_)) != cudaSuccess) {
    error_ = std::string("stream synchronize failed: ") + cudaGetErrorString(cudaGetLastError());
    return kTfLiteError;
  }
  return kTfLiteOk;
}

TfLiteStatus Graph::ReadTensorAsync(int i, void* dst, size_t bytes) {
  TfLiteTensor& t = tensors_[i];
  if (bytes > t.bytes || !t.data.raw || !device_arena_) {
    error_ = "ReadTensorAsync: size mismatch, unallocated tensor or host arena";
    return kTfLiteError;
  }
  if (cudaMemcpyAsync(dst, t.data.raw, bytes, cudaMemcpyDefault,
                      static_cast<cudaStream_t>(stream_)) != cudaSuccess) {
    error_ = "ReadTensorAsync: copy failed";
    return kTfLiteError;
  }
  return kTfLiteOk;
}

TfLiteStatus Graph::WriteTensor(int i, const void* src, size_t bytes) {
  TfLiteTensor& t = tensors_[i];
  if (bytes > t.bytes || !t.data.raw) {
    error_ = "WriteTensor: size mismatch or unallocated tensor";
    return kTfLiteError;
  }
  if (device_arena_ && t.allocation_type == kTfLiteArenaRw) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (cudaMemcpyAsync(t.data.raw, src, bytes, cudaMemcpyDefault, s) != cudaSuccess) {
      error_ = "WriteTensor: H2D copy failed";
      return kTfLiteError;
    }
  } else {
    memcpy(t.data.raw, src, bytes);
  }
  return kTfLiteOk;
}

TfLiteStatus Graph::ReadTensor(int i, void* dst, size_t bytes) {
  TfLiteTensor& t = tensors_[i];
  if (bytes > t.bytes || !t.data.raw) {
    error_ = "ReadTensor: size mismatch or unallocated tensor";
    return kTfLiteError;
  }
  if (device_arena_ && t.allocation_type == kTfLiteArenaRw) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (cudaMemcpyAsync(dst, t.data.raw, bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      error_ = std::string("ReadTensor: D2H copy failed: ") +
               cudaGetErrorString(cudaGetLastError());
      return kTfLiteError;
    }
  } else {
    memcpy(dst, const_host_[i].empty() ? t.data.raw : reinterpret_cast<char*>(const_host_[i].data()),
           bytes);
  }
  return kTfLiteOk;
}

}  // namespace lce_b200
