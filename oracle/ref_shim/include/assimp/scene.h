#pragma once
#include "Importer.hpp"
