#pragma once
#include "lazyten.hh"
