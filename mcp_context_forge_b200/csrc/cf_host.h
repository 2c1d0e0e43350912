// cf_host.h — internal definition of cf_builder (shared by cf_host.cpp, cfgpu.cu and the
// CPU-test helper).
#pragma once
#include <new>
#include <string>
#include <vector>

#include "../../include/cfgpu.h"
#include "re_backend.h"

struct cf_builder {
  std::vector<cfre::PatternIn> pats;
  std::vector<uint8_t> ordered;
  std::vector<std::vector<uint8_t>> repl;
  std::vector<uint8_t> has_repl;
  std::vector<std::vector<uint32_t>> tmpl;   // per pattern: {kind, a, b} triples when the replacement references groups, else empty
  cfre::CharSet word;
  cfre::CompileOut out;
  bool compiled = false;
  std::string err;
};
