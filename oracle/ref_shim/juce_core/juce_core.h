// TEST INFRASTRUCTURE ONLY: stand-in so that reference headers which include <juce_core/juce_core.h>
// (src/UltraHighRateDCBlocker.h) compile against the same shim as JuceHeader.h.
#pragma once
#include "../JuceHeader.h"
