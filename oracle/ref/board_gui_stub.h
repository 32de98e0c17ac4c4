#pragma once
/* TEST INFRASTRUCTURE — stands in for src/risk_game/board/board_gui.h (Windows-only ImGui / Direct3D viewer, compiled by the
   reference's CMake only under WIN32 AND GUI, CMakeLists.txt:55-64) so that src/alphazero_risk.h compiles on Linux. */
#include "../game/game.h"
