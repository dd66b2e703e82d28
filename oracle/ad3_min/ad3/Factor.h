// TEST INFRASTRUCTURE ONLY (oracle). budget.h:5 includes "ad3/Factor.h"; the
// only thing the reference's factors need from it is already declared by the
// minimal GenericFactor restatement.
#pragma once
#include "ad3/GenericFactor.h"
