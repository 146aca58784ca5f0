// Test infrastructure (oracle/): included by ProposalEngine.cpp:34 but never used.
#pragma once
