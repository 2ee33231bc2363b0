/*
 * core.h -- the ABI kit of the gpusdrpipeline boundary, re-authored for g++/nvcc.
 *
 * Mirrors (same names, same layouts, same virtual-function order -- binaries built against either set of headers
 * interoperate) what the reference spreads over include/gpusdrpipeline/{Status,GSDefs,IRef,GSLog,Result,GSErrors,
 * CudaErrors,SampleType,Modulation,am,fm}.h.  The per-file names of the reference still exist next to this file as
 * one-line forwarding headers, so `#include <gpusdrpipeline/Result.h>` keeps working.
 *
 *   Status           reference Status.h:22-34      (uint32_t, ten codes, fixed order)
 *   IRef             reference IRef.h:30-38        (ref/unref, protected virtual dtor)
 *   Result<T>        reference Result.h:28-52      (pack(8) {status, value}; RefResult for IRef types, ValResult else)
 *   floating refs    reference IRef.h:282-298      (objects are born with count 0; unref at <= 1 deletes)
 *
 * Difference from the reference that does NOT change the ABI: `#pragma pack` sits outside the template declarations
 * (the reference puts it between `template <...>` and `struct`, which only clang parses).
 */
#ifndef GPUSDRPIPELINE_ABI_CORE_H
#define GPUSDRPIPELINE_ABI_CORE_H

#include <cuda_runtime.h>
#include <gpusdrpipeline/gpusdrpipeline_export.h>

#include <atomic>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

/* ---- linkage / declaration helpers (reference GSDefs.h:22-62) -------------------------------------------- */
#define GS_C_LINKAGE extern "C"
#define GS_EXPORT GS_C_LINKAGE GS_PUBLIC
#define GS_FMT_STR(p) p
#ifdef __GNUC__
#define GS_FMT_ATTR(FMT_OFFSET, PARAM_OFFSET) __attribute__((format(printf, FMT_OFFSET, PARAM_OFFSET)))
#else
#define GS_FMT_ATTR(FMT_OFFSET, PARAM_OFFSET)
#endif

// abstract interface deriving from IRef: protected default ctor, protected virtual dtor
#define ABSTRACT_IREF(CLASS_NAME__)  \
 protected:                          \
  CLASS_NAME__() noexcept = default; \
  ~CLASS_NAME__() override = default;

// concrete ref-counted class: ref()/unref() through an embedded RefCt, deletion only from inside the library
#define REF_COUNTED_NO_DESTRUCTOR(CLASS_TYPE__)                                \
 public:                                                                       \
  void ref() const noexcept final { mRefCt.ref(); }                            \
  void unref() const noexcept final { mRefCt.unref(); }                        \
                                                                               \
 private:                                                                      \
  static void mSelfDeleter(CLASS_TYPE__* selfPtr) noexcept { delete selfPtr; } \
  RefCt<CLASS_TYPE__> mRefCt { this, mSelfDeleter }

#define REF_COUNTED(REF_CT_CLASS_TYPE__)  \
 private:                                 \
  ~REF_CT_CLASS_TYPE__() final = default; \
  REF_COUNTED_NO_DESTRUCTOR(REF_CT_CLASS_TYPE__)

/* ---- Status (reference Status.h:22-78) ------------------------------------------------------------------- */
using Status = uint32_t;
enum Status_ {
  Status_Success,
  Status_UnknownError,
  Status_OutOfMemory,
  Status_RuntimeError,
  Status_InvalidArgument,
  Status_InvalidState,
  Status_OutOfRange,
  Status_TimedOut,
  Status_NotFound,
  Status_ParseError,
};

// caller-side only: never lets an exception cross the library boundary
inline void throwIfError(Status status) {
  switch (status) {
    case Status_Success: return;
    case Status_OutOfMemory: throw std::bad_alloc();
    case Status_InvalidArgument: throw std::invalid_argument("Invalid Argument");
    case Status_OutOfRange: throw std::out_of_range("Out of Range");
    case Status_UnknownError: throw std::runtime_error("Unknown Error");
    case Status_RuntimeError: throw std::runtime_error("Error");
    case Status_InvalidState: throw std::runtime_error("Invalid State");
    case Status_TimedOut: throw std::runtime_error("Timed Out");
    case Status_NotFound: throw std::runtime_error("Not Found");
    case Status_ParseError: throw std::runtime_error("Parse Error");
    default: throw std::runtime_error("Error type [" + std::to_string(status) + "]");
  }
}

/* ---- sample / modulation enums (reference SampleType.h:20-25, Modulation.h:22-26) and band constants ------ */
using SampleType = uint32_t;
enum SampleType_ { SampleType_FloatComplex, SampleType_Float, SampleType_Int8Complex };
using Modulation = uint32_t;
enum Modulation_ { Modulation_Am, Modulation_Fm };
constexpr double kAmChannelBandwidth = 10e3;     // reference am.h:20
constexpr double kTauEu = 50e-6;                 // reference fm.h:20-27
constexpr double kTauUs = 75e-6;
constexpr double kNbfmChannelWidth = 15e3;
constexpr double kWbfmChannelWidth = 200e3;
constexpr double kNbfmFrequencyDeviation = 5e3;
constexpr double kWbfmFrequencyDeviation = 75e3;

/* ---- IRef and the smart pointers (reference IRef.h) ------------------------------------------------------ */
class IRef {
 public:
  virtual void ref() const noexcept = 0;
  virtual void unref() const noexcept = 0;

 protected:
  IRef() noexcept = default;
  virtual ~IRef() = default;
};

// holds one reference for its whole life; not re-assignable
template <typename T>
class ImmutableRef final {
 public:
  ImmutableRef() noexcept : mPtr(nullptr) {}
  ImmutableRef(T* p) noexcept : mPtr(p) { acquire(); }
  ImmutableRef(const ImmutableRef& o) noexcept : mPtr(o.mPtr) { acquire(); }
  ImmutableRef(ImmutableRef&& o) noexcept : mPtr(o.mPtr) { acquire(); }
  ~ImmutableRef() noexcept {
    if (mPtr) mPtr->unref();
  }
  ImmutableRef& operator=(const ImmutableRef&) = delete;
  ImmutableRef& operator=(ImmutableRef&&) = delete;

  operator T*() const noexcept { return mPtr; }
  T* operator->() const noexcept { return mPtr; }
  T* get() const noexcept { return mPtr; }
  bool operator==(const ImmutableRef& o) const noexcept { return mPtr == o.mPtr; }
  bool operator!=(const ImmutableRef& o) const noexcept { return mPtr != o.mPtr; }
  bool operator==(T* o) const noexcept { return mPtr == o; }
  bool operator!=(T* o) const noexcept { return mPtr != o; }

 private:
  void acquire() const noexcept {
    if (mPtr) mPtr->ref();
  }
  T* const mPtr;
};

template <typename T>
using ConstRef = const ImmutableRef<T>;

// re-assignable reference
template <typename T, typename = typename std::enable_if<std::is_base_of<IRef, T>::value>::type>
class Ref final {
 public:
  Ref() noexcept : mPtr(nullptr) {}
  Ref(T* p) noexcept : mPtr(nullptr) { reset(p); }
  Ref(const Ref& o) noexcept : mPtr(nullptr) { reset(o.mPtr.load()); }
  Ref(Ref&& o) noexcept : mPtr(o.mPtr.exchange(nullptr)) {}
  Ref(const ImmutableRef<T>& o) noexcept : mPtr(nullptr) { reset(o.get()); }
  ~Ref() noexcept { reset(nullptr); }

  Ref& operator=(T* p) noexcept {
    reset(p);
    return *this;
  }
  Ref& operator=(const Ref& o) noexcept {
    if (&o != this) reset(o.mPtr.load());
    return *this;
  }
  Ref& operator=(const ImmutableRef<T>& o) noexcept {
    reset(o.get());
    return *this;
  }
  Ref& operator=(Ref&& o) noexcept {
    if (&o != this) {
      T* old = mPtr.exchange(o.mPtr.exchange(nullptr));
      if (old) old->unref();
    }
    return *this;
  }
  void reset() noexcept { reset(nullptr); }
  void reset(T* p) noexcept {
    if (p) p->ref();  // before dropping the old one: p may be the object already held
    T* old = mPtr.exchange(p);
    if (old) old->unref();
  }

  operator ImmutableRef<T>() const noexcept { return ImmutableRef<T>(mPtr.load()); }
  ImmutableRef<T> operator->() const noexcept { return ImmutableRef<T>(mPtr.load()); }
  ImmutableRef<T> get() const noexcept { return ImmutableRef<T>(mPtr.load()); }
  bool operator==(const Ref& o) const noexcept { return mPtr.load() == o.mPtr.load(); }
  bool operator!=(const Ref& o) const noexcept { return mPtr.load() != o.mPtr.load(); }
  bool operator==(T* o) const noexcept { return mPtr.load() == o; }
  bool operator!=(T* o) const noexcept { return mPtr.load() != o; }

 private:
  std::atomic<T*> mPtr;
};

// holds a reference that can be handed out raw (for `return {Status_Success, ref.steal()}`)
template <typename T>
class StealableRef final {
 public:
  StealableRef() noexcept : mPtr(nullptr) {}
  explicit StealableRef(T* p) noexcept : mPtr(p) {
    if (p) p->ref();
  }
  ~StealableRef() {
    if (T* p = steal()) p->unref();
  }
  StealableRef& operator=(T* p) {
    if (p) p->ref();
    if (T* old = mPtr.exchange(p)) old->unref();
    return *this;
  }
  T* steal() noexcept { return mPtr.exchange(nullptr); }
  operator ImmutableRef<T>() const noexcept { return ImmutableRef<T>(mPtr.load()); }
  ImmutableRef<T> operator->() const noexcept { return ImmutableRef<T>(mPtr.load()); }
  ImmutableRef<T> get() const noexcept { return ImmutableRef<T>(mPtr.load()); }

 private:
  std::atomic<T*> mPtr;
};

// embedded counter of a concrete class.  Starts at 0 ("floating"): the first holder's ref() makes it 1; unref() at
// <= 1 -- including on an object nobody ever reffed -- destroys it.  Layout: {count, context, callback}.
template <class T>
class RefCt final {
 public:
  RefCt(T* context, void (*onZero)(T*) noexcept) noexcept : mContext(context), mOnZero(onZero) {}
  RefCt(const RefCt&) = delete;
  RefCt& operator=(const RefCt&) = delete;
  void ref() const noexcept { mCount.fetch_add(1); }
  void unref() const noexcept {
    size_t before = mCount.load();
    while (before != 0 && !mCount.compare_exchange_weak(before, before - 1)) {
    }
    if (before <= 1) mOnZero(mContext);
  }

 private:
  mutable std::atomic_size_t mCount {0};
  T* const mContext;
  void (*const mOnZero)(T*) noexcept;
};

/* ---- logging (reference GSLog.h:27-56) ---------------------------------------------------------------- */
using LogLevel = uint32_t;
enum LogLevel_ { GSLOG_TRACE, GSLOG_DEBUG, GSLOG_INFO, GSLOG_WARN, GSLOG_ERROR, GSLOG_FATAL };

class ILogger : public virtual IRef {
 public:
  virtual void log(LogLevel level, const char* msgFmt, va_list args) noexcept = 0;
  ABSTRACT_IREF(ILogger);
};

GS_EXPORT [[nodiscard]] const char* gslogLevelName(LogLevel level) noexcept;
GS_EXPORT void gsvlog(LogLevel level, const char* fmt, va_list args) noexcept;
GS_EXPORT void gslogSetLogger(ILogger* logger) noexcept;
GS_EXPORT void gslogSetVerbosity(LogLevel level) noexcept;
GS_EXPORT GS_FMT_ATTR(1, 2) void gslogt(const char* fmt, ...) noexcept;
GS_EXPORT GS_FMT_ATTR(1, 2) void gslogd(const char* fmt, ...) noexcept;
GS_EXPORT GS_FMT_ATTR(1, 2) void gslogi(const char* fmt, ...) noexcept;
GS_EXPORT GS_FMT_ATTR(1, 2) void gslogw(const char* fmt, ...) noexcept;
GS_EXPORT GS_FMT_ATTR(1, 2) void gsloge(const char* fmt, ...) noexcept;
GS_EXPORT [[noreturn]] GS_FMT_ATTR(1, 2) void gslogf(const char* fmt, ...) noexcept;

/* ---- Result (reference Result.h:28-52) ------------------------------------------------------------------- */
#pragma pack(push, 8)
template <typename T>
struct RefResult {
  using ValueType = T*;
  const Status status;
  T* const value;
};
template <typename T>
struct ValResult {
  using ValueType = T;
  Status status;
  T value;
};
#pragma pack(pop)

template <typename T>
using Result = typename std::conditional<std::is_base_of<IRef, T>::value, RefResult<T>, ValResult<T>>::type;

// caller side: take the (floating) reference out of a result, throwing on error
template <typename T>
ImmutableRef<T> unwrap(RefResult<T>&& result) {
  ImmutableRef<T> value = result.value;
  throwIfError(result.status);
  return value;
}
template <typename T>
T* unwrapRaw(RefResult<T>&& result) {
  if (result.status != Status_Success && result.value != nullptr) result.value->unref();
  throwIfError(result.status);
  return result.value;
}
template <typename T>
T unwrap(ValResult<T>&& result) {
  throwIfError(result.status);
  return result.value;
}
template <typename Out, typename In>
inline RefResult<Out> ResultCast(const RefResult<In>& result) noexcept {
  return {result.status, result.value};
}
template <typename Out, typename In>
inline ValResult<Out> ResultCast(const ValResult<In>& result) noexcept {
  return {result.status, static_cast<Out>(result.value)};
}
template <typename T>
RefResult<T> makeRefResultNonNull(T* reffed) noexcept {
  return {reffed != nullptr ? Status_Success : Status_OutOfMemory, reffed};
}
template <typename T>
RefResult<T> makeRefResultNonNull(const ImmutableRef<T>& reffed) noexcept {
  return {reffed != nullptr ? Status_Success : Status_OutOfMemory, reffed.get()};
}
template <typename T>
Result<T> makeRefResultNullable(T* reffed) noexcept {
  return {Status_Success, reffed};
}
template <typename T>
Result<T> makeValResult(T value) noexcept {
  return {Status_Success, value};
}
template <typename T>
Result<T> errResult(Status status) {
  return {status, {}};
}
#define ERR_RESULT(errResultStatus__) \
  { errResultStatus__, {} }

/* status / result plumbing macros (reference Result.h:136-349): same names, same control flow */
#define GS_LOG_ERR_AT__(what__, code__) gsloge("Error [%d] in [%s] at %s:%d", static_cast<int>(code__), what__, __FILE__, __LINE__)

#define UNWRAP_OR_FWD_RESULT(assignValueToVar__, unwrapCmd__)   \
  do {                                                          \
    auto result__ = unwrapCmd__;                                \
    if (result__.status != Status_Success) {                    \
      GS_LOG_ERR_AT__(#unwrapCmd__, result__.status);           \
      return {result__.status, {}};                             \
    }                                                           \
    assignValueToVar__ = result__.value;                        \
  } while (false)
#define UNWRAP_MOVE_OR_FWD_RESULT(assignValueToVar__, unwrapCmd__) \
  do {                                                             \
    auto result__ = unwrapCmd__;                                   \
    if (result__.status != Status_Success) {                       \
      GS_LOG_ERR_AT__(#unwrapCmd__, result__.status);              \
      return {result__.status, {}};                                \
    }                                                              \
    assignValueToVar__ = std::move(result__.value);                \
  } while (false)
#define UNWRAP_OR_FWD_STATUS(assignValueToVar__, unwrapCmd__) \
  do {                                                        \
    auto result__ = unwrapCmd__;                              \
    if (result__.status != Status_Success) {                  \
      GS_LOG_ERR_AT__(#unwrapCmd__, result__.status);         \
      return result__.status;                                 \
    }                                                         \
    assignValueToVar__ = result__.value;                      \
  } while (false)
#define UNWRAP_OR_RETURN(assignValueToVar__, unwrapCmd__, retOnError__) \
  do {                                                                  \
    auto result__ = unwrapCmd__;                                        \
    if (result__.status != Status_Success) {                            \
      GS_LOG_ERR_AT__(#unwrapCmd__, result__.status);                   \
      return retOnError__;                                              \
    }                                                                   \
    assignValueToVar__ = result__.value;                                \
  } while (false)
#define DO_OR_FWD_ERR(unwrapCmd__)                    \
  do {                                                \
    auto result__ = unwrapCmd__;                      \
    if (result__.status != Status_Success) {          \
      GS_LOG_ERR_AT__(#unwrapCmd__, result__.status); \
      return {result__.status, {}};                   \
    }                                                 \
  } while (false)
#define WARN_IF_ERR(cmdReturningStatus__)                                                   \
  do {                                                                                      \
    const Status status__ = cmdReturningStatus__;                                           \
    if (status__ != Status_Success) GS_LOG_ERR_AT__(#cmdReturningStatus__, status__);       \
  } while (false)
#define FWD_IF_ERR(cmdReturningStatus__)                  \
  do {                                                    \
    const Status status__ = cmdReturningStatus__;         \
    if (status__ != Status_Success) {                     \
      GS_LOG_ERR_AT__(#cmdReturningStatus__, status__);   \
      return status__;                                    \
    }                                                     \
  } while (false)
#define THROW_IF_ERR(cmdReturningStatus__)                \
  do {                                                    \
    const Status status__ = cmdReturningStatus__;         \
    if (status__ != Status_Success) {                     \
      GS_LOG_ERR_AT__(#cmdReturningStatus__, status__);   \
      throwIfError(status__);                             \
    }                                                     \
  } while (false)
#define RET_IF_ERR(cmdReturningStatus__, returnValueOnErr__) \
  do {                                                       \
    const Status status__ = cmdReturningStatus__;            \
    if (status__ != Status_Success) {                        \
      GS_LOG_ERR_AT__(#cmdReturningStatus__, status__);      \
      return returnValueOnErr__;                             \
    }                                                        \
  } while (false)
#define FWD_IN_RESULT_IF_ERR(cmdReturningStatus__)        \
  do {                                                    \
    const Status status__ = cmdReturningStatus__;         \
    if (status__ != Status_Success) {                     \
      GS_LOG_ERR_AT__(#cmdReturningStatus__, status__);   \
      return {status__, {}};                              \
    }                                                     \
  } while (false)
#define NON_NULL_OR_RET(ptr__)                                            \
  do {                                                                    \
    if ((ptr__) == nullptr) {                                             \
      gsloge("%s cannot be null - at %s:%d", #ptr__, __FILE__, __LINE__); \
      return ERR_RESULT(Status_OutOfMemory);                              \
    }                                                                     \
  } while (false)
#define NON_NULL_PARAM_OR_RET(ptr__)                                      \
  do {                                                                    \
    if ((ptr__) == nullptr) {                                             \
      gsloge("%s cannot be null - at %s:%d", #ptr__, __FILE__, __LINE__); \
      return ERR_RESULT(Status_InvalidArgument);                          \
    }                                                                     \
  } while (false)

// exception -> Status at the library edge (reference Result.h:309-349)
#define GS_CATCH_TO__(wrap__)                                           \
  catch (const std::bad_alloc&) { return wrap__(Status_OutOfMemory); }  \
  catch (const std::out_of_range&) { return wrap__(Status_OutOfRange); } \
  catch (const std::invalid_argument&) { return wrap__(Status_InvalidArgument); } \
  catch (const std::runtime_error&) { return wrap__(Status_RuntimeError); } \
  catch (...) { return wrap__(Status_UnknownError); }
#define GS_STATUS_IDENTITY__(s__) s__
#define IF_CATCH_RETURN_STATUS GS_CATCH_TO__(GS_STATUS_IDENTITY__)
#define IF_CATCH_RETURN_RESULT GS_CATCH_TO__(ERR_RESULT)
#define DO_OR_RET_STATUS(doCmd__) \
  do {                            \
    try {                         \
      doCmd__;                    \
    }                             \
    IF_CATCH_RETURN_STATUS        \
  } while (false)
#define DO_OR_RET_ERR_RESULT(doCmd__) \
  do {                                \
    try {                             \
      doCmd__;                        \
    }                                 \
    IF_CATCH_RETURN_RESULT            \
  } while (false)

/* ---- requirement macros (reference GSErrors.h:28-213; the subset applications and nodes use) ------------- */
#define SSTREAM(x) static_cast<std::ostringstream&&>(std::ostringstream() << x).str()
#define GS_FAIL(x)                                                             \
  do {                                                                         \
    gslogf("%s - at %s:%d", SSTREAM(x).c_str(), __FILE__, __LINE__);           \
  } while (false)
#define GS_REQUIRE_OR_ABORT(requireCmd__, messageOnFalse__)                                                        \
  do {                                                                                                             \
    if (!(requireCmd__)) {                                                                                         \
      gsloge("Expression must be true [%s] - %s - at %s:%d", #requireCmd__, messageOnFalse__, __FILE__, __LINE__); \
      abort();                                                                                                     \
    }                                                                                                              \
  } while (false)
#define GS_REQUIRE_OR_RET(requireCmd__, messageOnFalse__, retValueOnFalse__)                                       \
  do {                                                                                                             \
    if (!(requireCmd__)) {                                                                                         \
      gsloge("Expression must be true [%s] - %s - at %s:%d", #requireCmd__, messageOnFalse__, __FILE__, __LINE__); \
      return retValueOnFalse__;                                                                                    \
    }                                                                                                              \
  } while (false)
#define GS_REQUIRE_OR_RET_STATUS(requireCmd__, messageOnFalse__) GS_REQUIRE_OR_RET(requireCmd__, messageOnFalse__, Status_InvalidArgument)
#define GS_REQUIRE_OR_RET_RESULT(requireCmd__, messageOnFalse__) GS_REQUIRE_OR_RET(requireCmd__, messageOnFalse__, ERR_RESULT(Status_InvalidArgument))
#define GS_REQUIRE_OR_RET_FMT(requireCmd__, returnOnFalse__, messageFmtOnFalse__, ...)      \
  do {                                                                                      \
    if (!(requireCmd__)) {                                                                  \
      gsloge("Expression must be true [%s] - at %s:%d", #requireCmd__, __FILE__, __LINE__); \
      gsloge(messageFmtOnFalse__, __VA_ARGS__);                                             \
      return returnOnFalse__;                                                               \
    }                                                                                       \
  } while (false)
#define GS_REQUIRE_OR_RET_STATUS_FMT(requireCmd__, ...) GS_REQUIRE_OR_RET_FMT(requireCmd__, Status_InvalidArgument, __VA_ARGS__)
#define GS_REQUIRE_OR_RET_RESULT_FMT(requireCmd__, ...) GS_REQUIRE_OR_RET_FMT(requireCmd__, ERR_RESULT(Status_InvalidArgument), __VA_ARGS__)
#define GS_REQUIRE_OR_THROW(requireCmd__, messageOnFalse__)                                                        \
  do {                                                                                                             \
    if (!(requireCmd__)) {                                                                                         \
      gsloge("Expression must be true [%s] - %s - at %s:%d", #requireCmd__, messageOnFalse__, __FILE__, __LINE__); \
      throw std::runtime_error(messageOnFalse__);                                                                  \
    }                                                                                                              \
  } while (false)

/* ---- CUDA error mapping (reference CudaErrors.h:25-44,117-198) ------------------------------------------- */
inline Status cudaErrorToStatus(cudaError_t cudaError) {
  switch (cudaError) {
    case cudaSuccess: return Status_Success;
    case cudaErrorInvalidValue: return Status_InvalidArgument;
    case cudaErrorIllegalAddress: return Status_OutOfRange;
    case cudaErrorIllegalState: return Status_InvalidState;
    case cudaErrorMemoryAllocation: return Status_OutOfMemory;
    case cudaErrorInvalidDevice:
    case cudaErrorFileNotFound:
    case cudaErrorJitCompilerNotFound:
    case cudaErrorSharedObjectSymbolNotFound: return Status_NotFound;
    default: return Status_RuntimeError;
  }
}
#define SAFE_CUDA_OR_RET(cudaCmd__, safeCudaRetOnFail__)                                                              \
  do {                                                                                                                \
    const cudaError_t safeCudaStatus__ = (cudaCmd__);                                                                 \
    if (safeCudaStatus__ != cudaSuccess) {                                                                            \
      gsloge("CUDA error %s: %s (%d). At %s:%d", #cudaCmd__, cudaGetErrorName(safeCudaStatus__), safeCudaStatus__, __FILE__, __LINE__); \
      return safeCudaRetOnFail__;                                                                                     \
    }                                                                                                                 \
  } while (false)
#define SAFE_CUDA_OR_RET_STATUS(cudaCmd__) SAFE_CUDA_OR_RET(cudaCmd__, cudaErrorToStatus(safeCudaStatus__))
#define SAFE_CUDA_OR_RET_RESULT(cudaCmd__) SAFE_CUDA_OR_RET(cudaCmd__, ERR_RESULT(cudaErrorToStatus(safeCudaStatus__)))
#define SAFE_CUDA_OR_THROW(cudaCmd__)                                                                                 \
  do {                                                                                                                \
    const cudaError_t safeCudaStatus__ = (cudaCmd__);                                                                 \
    if (safeCudaStatus__ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorName(safeCudaStatus__)); \
  } while (false)
#define SAFE_CUDA_WARN_ONLY(cudaCmd__)                                                                                \
  do {                                                                                                                \
    const cudaError_t safeCudaStatus__ = (cudaCmd__);                                                                 \
    if (safeCudaStatus__ != cudaSuccess) gslogw("CUDA error %s: %s. At %s:%d", #cudaCmd__, cudaGetErrorName(safeCudaStatus__), __FILE__, __LINE__); \
  } while (false)

/* ---- current-device helpers (reference util/CudaUtil.h:25, util/CudaDevicePushPop.h:27-79) ---------------- */
GS_EXPORT [[nodiscard]] Result<int32_t> gsGetCurrentCudaDevice() noexcept;

class CudaDevicePushPop final {
 public:
  explicit CudaDevicePushPop(int32_t device) noexcept {
    if (cudaGetDevice(&mPrevious) != cudaSuccess) mPrevious = -1;
    mStatus = mPrevious == device ? cudaSuccess : cudaSetDevice(device);
    mChanged = mStatus == cudaSuccess && mPrevious != device;
  }
  ~CudaDevicePushPop() noexcept {
    if (mChanged && mPrevious >= 0) cudaSetDevice(mPrevious);
  }
  [[nodiscard]] cudaError_t status() const noexcept { return mStatus; }

 private:
  int mPrevious = -1;
  bool mChanged = false;
  cudaError_t mStatus = cudaSuccess;
};
#define CUDA_DEV_PUSH_POP_OR_RET(deviceIndex__, returnOnFailure__)       \
  CudaDevicePushPop cudaDevicePushPop__(deviceIndex__);                  \
  if (cudaDevicePushPop__.status() != cudaSuccess) return returnOnFailure__
#define CUDA_DEV_PUSH_POP_OR_RET_STATUS(deviceIndex__) CUDA_DEV_PUSH_POP_OR_RET(deviceIndex__, cudaErrorToStatus(cudaDevicePushPop__.status()))
#define CUDA_DEV_PUSH_POP_OR_RET_RESULT(deviceIndex__) CUDA_DEV_PUSH_POP_OR_RET(deviceIndex__, ERR_RESULT(cudaErrorToStatus(cudaDevicePushPop__.status())))
#define CUDA_DEV_PUSH_POP_OR_THROW(deviceIndex__)                        \
  CudaDevicePushPop cudaDevicePushPop__(deviceIndex__);                  \
  if (cudaDevicePushPop__.status() != cudaSuccess) throw std::runtime_error("cudaSetDevice failed")

#endif  // GPUSDRPIPELINE_ABI_CORE_H
