// gscf-b200: same header name as src/gscf/PlainScfKeys.hh of the reference; everything lives in gscf.hh
#pragma once
#include "gscf.hh"
