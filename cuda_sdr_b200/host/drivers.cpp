// Graph drivers: SteppingDriver (one pull-based pass per doFilter), FilterDriver (a sub-graph presented as one Filter),
// the JSON "Component" builder and the Graphviz dump.  Behaviour follows reference src/driver/SteppingDriver.cpp:193-366,
// FilterDriver.cpp:163-214, FilterDriverFactory.cpp:27-179, DriverToDot.cpp.
//
// Compared with the reference's scheduler the per-edge bookkeeping is flat: edges live in two vectors indexed once at
// connect() time, so a step does no map lookups, no string copies and no allocations beyond the small per-call vectors.
#include <algorithm>
#include <cstring>
#include <map>
#include <string>

#include "internal.h"
#include "json_min.h"

namespace gs {
namespace {

struct Edge {
  Ref<Source> source;
  size_t sourcePort;
  Ref<Sink> sink;
  size_t sinkPort;
};

// the graph bookkeeping shared by both drivers
class Graph {
 public:
  Status connect(Source* source, size_t sourcePort, Sink* sink, size_t sinkPort) noexcept {
    GS_REQUIRE_OR_RET_STATUS(source != nullptr && sink != nullptr, "connect() needs a source and a sink");
    try {
      mEdges.push_back({Ref<Source>(source), sourcePort, Ref<Sink>(sink), sinkPort});
      addNode(source);
      addNode(sink);
      return Status_Success;
    }
    IF_CATCH_RETURN_STATUS
  }
  Status setupNode(Node* node, const char* name) noexcept {
    GS_REQUIRE_OR_RET_STATUS(node != nullptr && name != nullptr, "setupNode() needs a node and a name");
    try {
      addNode(node);
      mNames[node] = name;
      return Status_Success;
    }
    IF_CATCH_RETURN_STATUS
  }
  std::string nameOf(Node* node) const {
    const auto it = mNames.find(node);
    if (it != mNames.end()) return it->second;
    char buf[32];
    snprintf(buf, sizeof(buf), "%p", static_cast<void*>(node));
    return buf;
  }
  bool hasName(Node* node) const { return mNames.find(node) != mNames.end(); }

  // sinks nothing in this graph reads from: the ends the pull starts at
  std::vector<Sink*> tails() const {
    std::vector<Sink*> out;
    for (const Edge& e : mEdges) {
      Sink* sink = e.sink.get().get();
      Source* asSource = sink->asSource();
      bool feedsSomething = false;
      if (asSource != nullptr)
        for (const Edge& o : mEdges) feedsSomething = feedsSomething || o.source.get().get() == asSource;
      if (!feedsSomething && std::find(out.begin(), out.end(), sink) == out.end()) out.push_back(sink);
    }
    return out;
  }
  size_t outputPortCount(Source* source) const {
    size_t n = 0;
    for (const Edge& e : mEdges)
      if (e.source.get().get() == source && e.sourcePort + 1 > n) n = e.sourcePort + 1;
    return n;
  }
  bool hasDataOnAllPorts(Source* source) const {
    const size_t ports = outputPortCount(source);
    for (size_t p = 0; p < ports; p++)
      if (source->getOutputDataSize(p) == 0) return false;
    return ports > 0;
  }

  Status pullInto(Sink* sink, int depth = 0) {
    GS_REQUIRE_OR_RET(depth < 256, "graph too deep (cycle?)", Status_InvalidState);
    for (const Edge& e : mEdges) {
      if (e.sink.get().get() != sink) continue;
      Source* source = e.source.get().get();
      Sink* upstream = source->asSink();
      if (upstream != nullptr && !hasDataOnAllPorts(source)) {
        FWD_IF_ERR(pullInto(upstream, depth + 1));
        if (!hasDataOnAllPorts(source)) return Status_Success;  // nothing came out this pass
      }
      if (hasDataOnAllPorts(source)) FWD_IF_ERR(pushFrom(source));
    }
    return Status_Success;
  }

  // one readOutput() of `source` straight into the input buffers of the sinks connected to it (zero copy for the first
  // sink of each port, a copy for every further one)
  Status pushFrom(Source* source) {
    const size_t ports = outputPortCount(source);
    std::vector<IBuffer*> outputs(ports, nullptr);
    std::vector<Ref<IBuffer>> keep;
    std::vector<std::pair<const Edge*, IBuffer*>> targets;
    for (const Edge& e : mEdges) {
      if (e.source.get().get() != source) continue;
      size_t alignment = source->getOutputSizeAlignment(e.sourcePort);
      if (alignment == 0) alignment = 1;
      const size_t preferred = e.sink.get()->preferredInputBufferSize(e.sinkPort);
      const size_t available = source->getOutputDataSize(e.sourcePort);
      const size_t smaller = preferred < available ? preferred : available;
      const size_t bytes = smaller > SIZE_MAX - alignment ? smaller / alignment * alignment : (smaller + alignment - 1) / alignment * alignment;
      Ref<IBuffer> buffer;
      UNWRAP_OR_FWD_STATUS(buffer, e.sink.get()->requestBuffer(e.sinkPort, bytes));
      keep.push_back(buffer);
      targets.emplace_back(&e, buffer.get().get());
      if (outputs[e.sourcePort] == nullptr) outputs[e.sourcePort] = buffer.get().get();
    }
    FWD_IF_ERR(source->readOutput(outputs.data(), outputs.size()));
    for (const auto& t : targets) {
      IBuffer* filled = outputs[t.first->sourcePort];
      const size_t bytes = filled->range()->used();
      if (t.second != filled) {
        IBufferCopier* copier = source->getOutputCopier(t.first->sourcePort);
        GS_REQUIRE_OR_RET(copier != nullptr, "a source with several sinks on one port needs an output copier", Status_InvalidState);
        FWD_IF_ERR(copier->copy(t.second->writePtr(), filled->readPtr(), bytes));
      }
      FWD_IF_ERR(t.first->sink.get()->commitBuffer(t.first->sinkPort, bytes));
    }
    return Status_Success;
  }

  const std::vector<Edge>& edges() const { return mEdges; }
  const std::vector<Ref<Node>>& nodes() const { return mNodes; }

 private:
  void addNode(Node* node) {
    for (const auto& n : mNodes)
      if (n.get().get() == node) return;
    mNodes.emplace_back(node);
  }
  std::vector<Edge> mEdges;
  std::vector<Ref<Node>> mNodes;
  std::map<Node*, std::string> mNames;
};

size_t copyName(const std::string& name, char* out, size_t outLen) {
  if (out != nullptr && outLen > 0) {
    const size_t n = name.size() < outLen - 1 ? name.size() : outLen - 1;
    memcpy(out, name.data(), n);
    out[n] = '\0';
  }
  return name.size() + 1;
}

#define GS_DRIVER_METHODS(self__)                                                                                                  \
  Status connect(Source* source, size_t sourcePort, Sink* sink, size_t sinkPort) noexcept final {                                   \
    return mGraph.connect(source, sourcePort, sink, sinkPort);                                                                     \
  }                                                                                                                                \
  Status setupNode(Node* node, const char* functionInGraph) noexcept final { return mGraph.setupNode(node, functionInGraph); }     \
  void iterateOverConnections(void* context, void (*it)(IDriver*, void*, Source*, size_t, Sink*, size_t) noexcept) noexcept final { \
    for (const Edge& e : mGraph.edges()) it(self__, context, e.source.get().get(), e.sourcePort, e.sink.get().get(), e.sinkPort);    \
  }                                                                                                                                \
  void iterateOverNodes(void* context, void (*it)(IDriver*, void*, Node*) noexcept) noexcept final {                                \
    for (const auto& n : mGraph.nodes()) it(self__, context, n.get().get());                                                        \
  }                                                                                                                                \
  void iterateOverNodeAttributes(Node* node, void* context, void (*it)(IDriver*, Node*, void*, const char*, const char*) noexcept) noexcept final { \
    if (mGraph.hasName(node)) it(self__, node, context, "name", mGraph.nameOf(node).c_str());                                       \
  }                                                                                                                                \
  size_t getNodeName(Node* node, char* name, size_t nameBufLen, bool* foundOut) noexcept final {                                    \
    if (foundOut != nullptr) *foundOut = mGraph.hasName(node);                                                                     \
    return copyName(mGraph.nameOf(node), name, nameBufLen);                                                                        \
  }

class SteppingDriver final : public ISteppingDriver {
 public:
  SteppingDriver() noexcept = default;
  GS_DRIVER_METHODS(this)
  Status doFilter() noexcept final {
    try {
      for (Sink* tail : mGraph.tails()) FWD_IF_ERR(mGraph.pullInto(tail));
      return Status_Success;
    }
    IF_CATCH_RETURN_STATUS
  }

 private:
  Graph mGraph;
  REF_COUNTED(SteppingDriver);
};

// A sub-graph with one designated input sink and one designated output source, usable wherever a Filter is.
class FilterDriver final : public IFilterDriver {
 public:
  FilterDriver() noexcept = default;
  GS_DRIVER_METHODS(this)
  void setDriverInput(Sink* sink) noexcept final { mInput = sink; }
  void setDriverOutput(Source* source) noexcept final { mOutput = source; }

  Result<IBuffer> requestBuffer(size_t port, size_t byteCount) noexcept final {
    GS_REQUIRE_OR_RET_RESULT(mInput != nullptr, "The driver's input node has not been set");
    return mInput.get()->requestBuffer(port, byteCount);
  }
  // committing feeds the inner graph and runs it right away (FilterDriver.cpp:163-171)
  Status commitBuffer(size_t port, size_t byteCount) noexcept final {
    GS_REQUIRE_OR_RET_STATUS(mInput != nullptr, "The driver's input node has not been set");
    FWD_IF_ERR(mInput.get()->commitBuffer(port, byteCount));
    return runInner();
  }
  size_t preferredInputBufferSize(size_t port) noexcept final { return mInput != nullptr ? mInput.get()->preferredInputBufferSize(port) : 0; }
  size_t getOutputDataSize(size_t port) noexcept final {
    if (mOutput == nullptr) return 0;
    if (mOutput.get()->getOutputDataSize(port) == 0) (void)runInner();
    return mOutput.get()->getOutputDataSize(port);
  }
  size_t getOutputSizeAlignment(size_t port) noexcept final { return mOutput != nullptr ? mOutput.get()->getOutputSizeAlignment(port) : 1; }
  IBufferCopier* getOutputCopier(size_t port) noexcept final { return mOutput != nullptr ? mOutput.get()->getOutputCopier(port) : nullptr; }
  Status readOutput(IBuffer** bufs, size_t numPorts) noexcept final {
    GS_REQUIRE_OR_RET_STATUS(mOutput != nullptr, "The driver's output node has not been set");
    if (mOutput.get()->getOutputDataSize(0) == 0) FWD_IF_ERR(runInner());
    return mOutput.get()->readOutput(bufs, numPorts);
  }

 private:
  Status runInner() noexcept {
    try {
      Sink* outSink = mOutput != nullptr ? mOutput.get()->asSink() : nullptr;
      if (outSink != nullptr) return mGraph.pullInto(outSink);
      for (Sink* tail : mGraph.tails()) FWD_IF_ERR(mGraph.pullInto(tail));
      return Status_Success;
    }
    IF_CATCH_RETURN_STATUS
  }
  Graph mGraph;
  Ref<Sink> mInput;
  Ref<Source> mOutput;
  REF_COUNTED(FilterDriver);
};

class SteppingFactory final : public ISteppingDriverFactory {
 public:
  SteppingFactory() noexcept = default;
  Result<ISteppingDriver> createSteppingDriver() noexcept final { return makeRefResultNonNull<ISteppingDriver>(new (std::nothrow) SteppingDriver()); }
  REF_COUNTED(SteppingFactory);
};

// JSON "Component" -- the reference's schema (FilterDriverFactory.cpp:27-179, example at :181-274):
//   {"nodes":       {"<id>": {"type": t, <the node's own parameters inline>}, ...},
//    "connections": [{"source": id, "sourcePort": p, "sink": id, "sinkPort": p}, ...],       (ports default to 0)
//    "inputPorts":  [{"exposedPort": n, "mapped": {"node": id, "port": p}}, ...],            -> a PortRemappingSink
//    "outputPorts": [{"exposedPort": n, "mapped": {"node": id, "port": p}}, ...]}            -> a PortRemappingSource
// Each node is created from ITS OWN definition (the reference passes the whole Component text to every node factory,
// FilterDriverFactory.cpp:51 -- a defect no node factory could work with; not reproduced).  Also accepted: "outputPort":
// id (the key RfToPcmAudioFactory.cpp:303 emits), "inputNode" / "outputNode": id, and "nodes" as an array of
// {"name", "type", "parameters"} (this library's first-round form).
class FilterDriverFactory final : public IFilterDriverFactory {
 public:
  explicit FilterDriverFactory(IFactories* f) noexcept : mFactories(f) {}
  Result<IFilterDriver> createFilterDriver() noexcept final { return makeRefResultNonNull<IFilterDriver>(new (std::nothrow) FilterDriver()); }
  Result<Node> create(const char* json) noexcept final {
    try {
      const Json spec = Json::parse(json);
      // raw and floating (count 0) until it is returned; every early exit below destroys it through the guard
      IFilterDriver* const driver = new (std::nothrow) FilterDriver();
      NON_NULL_OR_RET(driver);
      struct Guard {
        IFilterDriver* d;
        ~Guard() {
          if (d != nullptr) d->unref();
        }
      } guard {driver};
      std::map<std::string, Ref<Node>> nodes;
      auto addNode = [&](const std::string& id, const Json& def, const std::string& params) -> Status {
        if (!def.contains("type")) {
          gsloge("Node definition for [%s] does not contain a type.", id.c_str());
          return Status_InvalidArgument;
        }
        if (nodes.count(id) != 0) {
          gsloge("Duplicate definition for node [%s].", id.c_str());
          return Status_InvalidArgument;
        }
        Ref<Node> node;
        UNWRAP_OR_FWD_STATUS(node, createNode(def.at("type").str().c_str(), params.c_str()));
        nodes.emplace(id, node);
        const std::string label = def.contains("description") && def.at("description").kind == Json::String ? def.at("description").str() : id;
        return driver->setupNode(node.get(), label.c_str());
      };
      const Json& nodeDefs = spec.at("nodes");
      if (nodeDefs.kind == Json::Object) {
        for (const auto& kv : nodeDefs.members) FWD_IN_RESULT_IF_ERR(addNode(kv.first, kv.second, kv.second.dump()));
      } else {
        for (const Json& n : nodeDefs.array())
          FWD_IN_RESULT_IF_ERR(addNode(n.at("name").str(), n, n.contains("parameters") ? n.at("parameters").dump() : n.dump()));
      }
      auto find = [&](const std::string& id, const char* what) -> Node* {
        const auto it = nodes.find(id);
        if (it == nodes.end()) {
          gsloge("Cannot %s node [%s] because it was not defined.", what, id.c_str());
          return nullptr;
        }
        return it->second.get().get();
      };
      auto port = [](const Json& obj, const char* key) -> size_t { return obj.contains(key) ? static_cast<size_t>(obj.at(key).num()) : 0; };

      // ---- exposed input ports -> PortRemappingSink ------------------------------------------------------------
      if (spec.contains("inputPorts") && !spec.at("inputPorts").array().empty()) {
        Ref<IPortRemappingSink> mapper;
        UNWRAP_OR_FWD_RESULT(mapper, mFactories->getPortRemappingSinkFactory()->create());
        for (const Json& m : spec.at("inputPorts").array()) {
          const Json& mapped = m.at("mapped");
          Node* node = find(mapped.at("node").str(), "add an input port mapping with");
          if (node == nullptr) return ERR_RESULT(Status_NotFound);
          GS_REQUIRE_OR_RET_RESULT_FMT(node->asSink() != nullptr, "Cannot add an input port mapping with node [%s] because it is not a sink.",
                                       mapped.at("node").str().c_str());
          mapper.get()->addPortMapping(port(m, "exposedPort"), node->asSink(), port(mapped, "port"));
        }
        driver->setDriverInput(mapper.get());
      } else if (spec.contains("inputNode")) {
        Node* node = find(spec.at("inputNode").str(), "use as the input");
        if (node == nullptr) return ERR_RESULT(Status_NotFound);
        GS_REQUIRE_OR_RET_RESULT(node->asSink() != nullptr, "the input node must be a Sink");
        driver->setDriverInput(node->asSink());
      }
      // ---- exposed output ports -> PortRemappingSource ---------------------------------------------------------
      if (spec.contains("outputPorts") && !spec.at("outputPorts").array().empty()) {
        Ref<IPortRemappingSource> mapper;
        UNWRAP_OR_FWD_RESULT(mapper, mFactories->getPortRemappingSourceFactory()->create());
        for (const Json& m : spec.at("outputPorts").array()) {
          const Json& mapped = m.at("mapped");
          Node* node = find(mapped.at("node").str(), "add an output port mapping with");
          if (node == nullptr) return ERR_RESULT(Status_NotFound);
          GS_REQUIRE_OR_RET_RESULT_FMT(node->asSource() != nullptr, "Cannot add an output port mapping with node [%s] because it is not a source.",
                                       mapped.at("node").str().c_str());
          mapper.get()->addPortMapping(port(m, "exposedPort"), node->asSource(), port(mapped, "port"));
        }
        driver->setDriverOutput(mapper.get());
      } else if (spec.contains("outputPort") || spec.contains("outputNode")) {
        Node* node = find(spec.at(spec.contains("outputPort") ? "outputPort" : "outputNode").str(), "use as the output");
        if (node == nullptr) return ERR_RESULT(Status_NotFound);
        GS_REQUIRE_OR_RET_RESULT(node->asSource() != nullptr, "the output node must be a Source");
        driver->setDriverOutput(node->asSource());
      }
      // ---- connections ------------------------------------------------------------------------------------------
      if (spec.contains("connections")) {
        for (const Json& c : spec.at("connections").array()) {
          Node* from = find(c.at("source").str(), "connect source");
          Node* to = find(c.at("sink").str(), "connect sink");
          if (from == nullptr || to == nullptr) return ERR_RESULT(Status_InvalidArgument);
          GS_REQUIRE_OR_RET_RESULT_FMT(from->asSource() != nullptr, "Cannot connect [%s]: it is not a source.", c.at("source").str().c_str());
          GS_REQUIRE_OR_RET_RESULT_FMT(to->asSink() != nullptr, "Cannot connect [%s]: it is not a sink.", c.at("sink").str().c_str());
          FWD_IN_RESULT_IF_ERR(driver->connect(from->asSource(), port(c, "sourcePort"), to->asSink(), port(c, "sinkPort")));
        }
      }
      guard.d = nullptr;
      return makeRefResultNonNull<Node>(static_cast<Node*>(driver));
    } catch (const std::invalid_argument& e) {
      gsloge("Bad Component description: %s", e.what());
      return ERR_RESULT(Status_ParseError);
    }
    IF_CATCH_RETURN_RESULT
  }

 private:
  IFactories* const mFactories;
  REF_COUNTED(FilterDriverFactory);
};

// Graphviz: one node per graph node (its name), one edge per connection labelled with the port numbers
class DriverToDot final : public IDriverToDiagram {
 public:
  DriverToDot() noexcept = default;
  Result<size_t> convertToDot(IDriver* driver, const char* name, char* out, size_t outLen) noexcept final {
    try {
      NON_NULL_PARAM_OR_RET(driver);
      std::string dot = std::string("digraph \"") + (name ? name : "graph") + "\" {\n";
      struct Ctx {
        std::string* dot;
      } ctx {&dot};
      driver->iterateOverNodes(&ctx, [](IDriver* d, void* c, Node* node) noexcept {
        char buf[256];
        bool found = false;
        d->getNodeName(node, buf, sizeof(buf), &found);
        char line[384];
        snprintf(line, sizeof(line), "  n%p [label=\"%s\"];\n", static_cast<void*>(node), buf);
        *static_cast<Ctx*>(c)->dot += line;
      });
      driver->iterateOverConnections(&ctx, [](IDriver*, void* c, Source* source, size_t sp, Sink* sink, size_t kp) noexcept {
        char line[256];
        snprintf(line, sizeof(line), "  n%p -> n%p [label=\"%zu:%zu\"];\n", static_cast<void*>(static_cast<Node*>(source)),
                 static_cast<void*>(static_cast<Node*>(sink)), sp, kp);
        *static_cast<Ctx*>(c)->dot += line;
      });
      dot += "}\n";
      return makeValResult<size_t>(copyName(dot, out, outLen));
    }
    IF_CATCH_RETURN_RESULT
  }
  REF_COUNTED(DriverToDot);
};
class DriverToDotFactory final : public IDriverToDiagramFactory {
 public:
  DriverToDotFactory() noexcept = default;
  Result<IDriverToDiagram> create() const final { return makeRefResultNonNull<IDriverToDiagram>(new (std::nothrow) DriverToDot()); }
  REF_COUNTED(DriverToDotFactory);
};

}  // namespace

ISteppingDriverFactory* newSteppingDriverFactory() noexcept { return new (std::nothrow) SteppingFactory(); }
IFilterDriverFactory* newFilterDriverFactory(IFactories* f) noexcept { return new (std::nothrow) FilterDriverFactory(f); }
IDriverToDiagramFactory* newDriverToDotFactory() noexcept { return new (std::nothrow) DriverToDotFactory(); }

}  // namespace gs
