/*
 * drivers.h -- graph drivers.  Mirrors reference include/gpusdrpipeline/driver/*.h.
 *   ISteppingDriver::doFilter()  one pull-based pass over the graph (src/driver/SteppingDriver.cpp:193-366)
 *   IFilterDriver                a driver that is itself a Filter, i.e. a sub-graph as one node (FilterDriver.cpp)
 */
#ifndef GPUSDRPIPELINE_ABI_DRIVERS_H
#define GPUSDRPIPELINE_ABI_DRIVERS_H

#include <gpusdrpipeline/abi/nodes.h>

// driver/IDriver.h:23-64
class IDriver : public virtual Node {
 public:
  IDriver* asDriver() noexcept override { return this; }
  [[nodiscard]] virtual Status connect(Source* source, size_t sourcePort, Sink* sink, size_t sinkPort) noexcept = 0;
  [[nodiscard]] virtual Status setupNode(Node* node, const char* functionInGraph) noexcept = 0;
  virtual void iterateOverConnections(
      void* context,
      void (*connectionIterator)(IDriver* driver, void* context, Source* source, size_t sourcePort, Sink* sink, size_t sinkPort) noexcept) noexcept = 0;
  virtual void iterateOverNodes(void* context, void (*nodeIterator)(IDriver* driver, void* context, Node* node) noexcept) noexcept = 0;
  virtual void iterateOverNodeAttributes(
      Node* node, void* context,
      void (*nodeAttrIterator)(IDriver* driver, Node* node, void* context, const char* attrName, const char* attrVal) noexcept) noexcept = 0;
  virtual size_t getNodeName(Node* node, char* name, size_t nameBufLen, bool* foundOut) noexcept = 0;
  ABSTRACT_IREF(IDriver);
};

// driver/ISteppingDriver.h:22-27, ISteppingDriverFactory.h:22-27
class ISteppingDriver : public IDriver {
 public:
  [[nodiscard]] virtual Status doFilter() noexcept = 0;
  ABSTRACT_IREF(ISteppingDriver);
};
class ISteppingDriverFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<ISteppingDriver> createSteppingDriver() noexcept = 0;
  ABSTRACT_IREF(ISteppingDriverFactory);
};

// driver/IFilterDriver.h:23-36, IFilterDriverFactory.h:23-28
class IFilterDriver : public IDriver, public Filter {
 public:
  virtual void setDriverInput(Sink* sink) noexcept = 0;
  virtual void setDriverOutput(Source* source) noexcept = 0;
  ABSTRACT_IREF(IFilterDriver);
};
class IFilterDriverFactory : public INodeFactory {
 public:
  [[nodiscard]] virtual Result<IFilterDriver> createFilterDriver() noexcept = 0;
  ABSTRACT_IREF(IFilterDriverFactory);
};

// driver/IDriverToDiagram.h:24-41, IDriverToDiagramFactory.h:21-26 -- Graphviz dump of a driver's graph
class IDriverToDiagram : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<size_t> convertToDot(IDriver* driver, const char* name, char* diagramBuffer, size_t diagramSize) noexcept = 0;
  ABSTRACT_IREF(IDriverToDiagram);
};
class IDriverToDiagramFactory : public virtual IRef {
 public:
  [[nodiscard]] virtual Result<IDriverToDiagram> create() const = 0;
  ABSTRACT_IREF(IDriverToDiagramFactory);
};

#endif  // GPUSDRPIPELINE_ABI_DRIVERS_H
