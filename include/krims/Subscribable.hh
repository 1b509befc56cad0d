#pragma once
#include "krims.hh"
