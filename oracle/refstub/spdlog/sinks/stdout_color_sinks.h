#pragma once
#include "../spdlog.h"
