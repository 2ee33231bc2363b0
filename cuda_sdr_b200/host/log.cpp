// gslog* exports (reference include/gpusdrpipeline/GSLog.h:44-56, src/GSLog.cpp): one process-wide logger and
// verbosity, both replaceable; the default logger writes "[LEVEL] message" lines to stderr.
#include <cstring>
#include <mutex>

#include "internal.h"

namespace {

std::mutex g_logMutex;
ILogger* g_logger = nullptr;  // holds one reference
std::atomic<LogLevel> g_verbosity {GSLOG_INFO};

void emit(LogLevel level, const char* fmt, va_list args) noexcept {
  if (level < g_verbosity.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_logMutex);
  if (g_logger != nullptr) {
    g_logger->log(level, fmt, args);
    return;
  }
  fprintf(stderr, "[%s] ", gslogLevelName(level));
  vfprintf(stderr, fmt, args);
  const size_t n = strlen(fmt);
  if (n == 0 || fmt[n - 1] != '\n') fputc('\n', stderr);
}

}  // namespace

#define GS_LOG_BODY(level__)    \
  va_list args;                 \
  va_start(args, fmt);          \
  emit(level__, fmt, args);     \
  va_end(args)

GS_EXPORT const char* gslogLevelName(LogLevel level) noexcept {
  static const char* const names[] = {"TRACE", "DEBUG", "INFO", "WARN", "ERROR", "FATAL"};
  return level <= GSLOG_FATAL ? names[level] : "UNKNOWN";
}
GS_EXPORT void gsvlog(LogLevel level, const char* fmt, va_list args) noexcept { emit(level, fmt, args); }
GS_EXPORT void gslogSetLogger(ILogger* logger) noexcept {
  std::lock_guard<std::mutex> lock(g_logMutex);
  if (logger != nullptr) logger->ref();
  if (g_logger != nullptr) g_logger->unref();
  g_logger = logger;
}
GS_EXPORT void gslogSetVerbosity(LogLevel level) noexcept { g_verbosity.store(level); }
GS_EXPORT void gslogt(const char* fmt, ...) noexcept { GS_LOG_BODY(GSLOG_TRACE); }
GS_EXPORT void gslogd(const char* fmt, ...) noexcept { GS_LOG_BODY(GSLOG_DEBUG); }
GS_EXPORT void gslogi(const char* fmt, ...) noexcept { GS_LOG_BODY(GSLOG_INFO); }
GS_EXPORT void gslogw(const char* fmt, ...) noexcept { GS_LOG_BODY(GSLOG_WARN); }
GS_EXPORT void gsloge(const char* fmt, ...) noexcept { GS_LOG_BODY(GSLOG_ERROR); }
GS_EXPORT void gslogf(const char* fmt, ...) noexcept {
  GS_LOG_BODY(GSLOG_FATAL);
  abort();
}

GS_EXPORT Result<int32_t> gsGetCurrentCudaDevice() noexcept {
  int device = -1;
  SAFE_CUDA_OR_RET_RESULT(cudaGetDevice(&device));
  return makeValResult<int32_t>(device);
}
