// ORACLE — TEST INFRASTRUCTURE ONLY.  See IpTNLP.hpp: nothing of IpoptApplication is used by armtd_NLP itself.
#pragma once
#include "IpTNLP.hpp"
