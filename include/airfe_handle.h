// Shared helper of the class surfaces: RAII owner of an airfe_ctx plus the weight-directory convention.
#ifndef AIRFE_HANDLE_H_
#define AIRFE_HANDLE_H_
#include <memory>
#include <string>
#include "airfe_c.h"

namespace airfe_cpp {
// The reference's configs name ONNX files inside a model directory (PLNetConfig::SetModelPath, include/read_configs.h:39-49).
// The B200 build reads the converted *.afw containers from that same directory; AIRFE_WEIGHTS_DIR overrides it.
std::string weights_dir_from(const std::string& onnx_path);
struct CtxDeleter { void operator()(airfe_ctx* c) const { airfe_destroy(c); } };
typedef std::unique_ptr<airfe_ctx, CtxDeleter> CtxPtr;
}  // namespace airfe_cpp
#endif
