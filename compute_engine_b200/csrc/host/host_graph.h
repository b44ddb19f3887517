// host_graph.h -- the caller of the hot path: a small graph host that plays the
// role TFLite's Subgraph plays for the reference (tensorflow/lite/core/subgraph.cc:
// OpInit :1271, OpPrepare :1302, OpInvoke :1368, ResizeInputTensor :1170,
// AllocateTensors) but keeps the tensor arena in HBM, so consecutive ops never
// round-trip through the host, and can replay the op sequence as one CUDA graph.
//
// It drives ANY TfLiteRegistration {init, free, prepare, invoke}: this library's
// CUDA ops (lce_ops.cc), the float builtins the three model families need
// (builtin_ops.cc), or -- in CPU tests -- the oracle-backed registrations built
// under oracle/ (arena then lives on the host; no CUDA call is made).
#ifndef LCE_B200_HOST_GRAPH_H_
#define LCE_B200_HOST_GRAPH_H_

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "lce_b200_tflite.h"

namespace lce_b200 {

// MutableOpResolver stand-in (tensorflow/lite/mutable_op_resolver.h:83): custom ops
// by name, builtins by schema BuiltinOperator code.
class OpResolver {
 public:
  void AddCustom(const char* name, const TfLiteRegistration* registration);
  void AddBuiltin(int builtin_code, const TfLiteRegistration* registration);
  const TfLiteRegistration* FindCustom(const std::string& name) const;
  const TfLiteRegistration* FindBuiltin(int builtin_code) const;

 private:
  std::map<std::string, const TfLiteRegistration*> custom_;
  std::map<int, const TfLiteRegistration*> builtin_;
};

struct NodeRecord {
  TfLiteNode node{};
  const TfLiteRegistration* registration = nullptr;
  std::vector<uint8_t> custom_options;
  std::vector<uint8_t> builtin_blob;  // builtin parameter struct (node.builtin_data)
  bool initialized = false;
  std::string name;
};

class Graph {
 public:
  // device_arena: activations in HBM (cudaMalloc) vs host memory (CPU tests).
  explicit Graph(bool device_arena);
  ~Graph();
  Graph(const Graph&) = delete;
  Graph& operator=(const Graph&) = delete;

  // const_data != nullptr makes a constant (kTfLiteMmapRo) tensor; its bytes are
  // copied (host copy always; device copy too when `const_on_device`).
  int AddTensor(TfLiteType type, const std::vector<int>& dims, const void* const_data,
                size_t const_bytes, bool has_quant, float scale, int zero_point,
                const std::string& name, bool const_on_device = false);
  int AddNode(const TfLiteRegistration* registration, const std::vector<int>& inputs,
              const std::vector<int>& outputs, const uint8_t* custom_options,
              size_t options_len, const void* builtin_data = nullptr,
              size_t builtin_bytes = 0, const std::string& name = "");
  void SetInputs(const std::vector<int>& t) { inputs_ = t; }
  void SetOutputs(const std::vector<int>& t) { outputs_ = t; }
  const std::vector<int>& inputs() const { return inputs_; }
  const std::vector<int>& outputs() const { return outputs_; }

  // Graph-level fusion of the converter's residual-block pattern
  //   LceBconv2d(float out y) -> ADD(y, shortcut) [-> LceQuantize of the sum]
  // into one node (the bconv epilogue adds the shortcut and emits the packed signs for
  // the next binary layer). Bit-identical results; y must have no other consumer.
  // Returns the number of (ADD, LceQuantize) nodes removed. Call before AllocateTensors.
  int FuseResidualBlocks();

  // Fusions among the float builtins around the binary path:
  //   MAX_POOL_2D(2x2, stride 1, VALID) -> DEPTHWISE_CONV_2D(3x3)  =>  one node
  //   CONV_2D(3x3 s2 -> 16) -> DEPTHWISE_CONV_2D(3x3 s2) -> CONV_2D(1x1 -> 64)  =>  one node (opt-in)
  //   CONV_2D -> LceQuantize  =>  the conv writes the packed signs as a second output
  // (QuickNet's anti-aliased down-sampling). Bit-identical. Returns the nodes removed.
  int FuseFloatGlue();
  int FuseConvQuantize();  // part of FuseFloatGlue: CONV_2D -> LceQuantize => conv with 2 outputs
  int FuseStem();  // part of FuseFloatGlue: [DEQUANTIZE +] stem conv + depthwise conv -> one node

  // init (first time) + prepare of every node in order, then arena allocation.
  TfLiteStatus AllocateTensors();
  TfLiteStatus ResizeInputTensor(int tensor, const std::vector<int>& dims);
  TfLiteStatus Invoke();
  // Capture the node sequence into a CUDA graph and replay it on later Invokes.
  TfLiteStatus EnableCudaGraph(bool on);
  // Debug aid (TFLite's `preserve_all_tensors`): no arena reuse, so every intermediate
  // tensor can be read back after Invoke.
  void set_preserve_all_tensors(bool on) { preserve_all_ = on; allocated_ = false; }

  // Per-node device timing (eager mode only): CUDA events on the graph's stream
  // around every node's invoke; times accumulate until ResetProfile().
  void EnableProfiling(bool on) { profiling_ = on; }
  void ResetProfile();
  double NodeTimeMs(size_t node);   // synchronises the stream
  TfLiteStatus Synchronize();

  TfLiteTensor* tensor(int i) { return &tensors_[i]; }
  size_t num_tensors() const { return tensors_.size(); }
  size_t num_nodes() const { return nodes_.size(); }
  const NodeRecord& node(size_t i) const { return *nodes_[i]; }
  TfLiteStatus WriteTensor(int i, const void* host_src, size_t bytes);
  TfLiteStatus ReadTensor(int i, void* host_dst, size_t bytes);
  // async on the graph's stream; pinned host or device destinations; no synchronise
  TfLiteStatus ReadTensorAsync(int i, void* dst, size_t bytes);
  const std::string& last_error() const { return error_; }
  void set_error(const std::string& e) { error_ = e; }
  bool device_arena() const { return device_arena_; }
  size_t arena_bytes() const { return arena_bytes_; }
  void* stream() const { return stream_; }

 private:
  static TfLiteStatus ResizeTensorCb(TfLiteContext*, TfLiteTensor*, TfLiteIntArray*);
  static void ReportErrorCb(TfLiteContext*, const char* msg, ...);
  static TfLiteStatus AddTensorsCb(TfLiteContext*, int, int*);
  static TfLiteTensor* GetTensorCb(const TfLiteContext*, int);
  static TfLiteExternalContext* GetExternalContextCb(TfLiteContext*, TfLiteExternalContextType);
  static void SetExternalContextCb(TfLiteContext*, TfLiteExternalContextType,
                                   TfLiteExternalContext*);
  TfLiteStatus PlanArena();
  void FreeArena();
  void RefreshContext();

  bool device_arena_;
  std::vector<TfLiteTensor> tensors_;
  std::vector<std::string> names_;
  std::vector<std::vector<uint8_t>> const_host_;  // host copies of constants
  std::vector<void*> const_dev_;                   // device copies (or nullptr)
  std::vector<std::unique_ptr<NodeRecord>> nodes_;
  std::vector<int> inputs_, outputs_;
  TfLiteContext ctx_{};
  TfLiteExternalContext* external_[kTfLiteMaxExternalContexts] = {};
  std::string error_;
  void* arena_ = nullptr;
  size_t arena_bytes_ = 0;
  bool allocated_ = false;
  void* stream_ = nullptr;       // cudaStream_t
  bool use_cuda_graph_ = false;
  bool preserve_all_ = false;
  bool profiling_ = false;
  std::vector<std::vector<std::pair<void*, void*>>> node_events_;  // pending (start, stop)
  std::vector<double> node_ms_;
  bool warmed_ = false;          // one eager Invoke has run since the last allocation
  void* graph_exec_ = nullptr;   // cudaGraphExec_t
};

}  // namespace lce_b200
#endif
