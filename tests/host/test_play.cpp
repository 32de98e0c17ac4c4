// Standalone test of alphazero_risk_b200/host/az_play.hpp with a mirror of the reference's GameResults (game/game.h:10-29).
// Without a GPU: construction must fail loudly -> NO_DEVICE_OK.  With a GPU: a 40-game match (--mcts=16, -t 2) -> PLAY_OK.
#include <cstdio>
#include <vector>
#include "az_play.hpp"

struct PlayerGameResult { int win = 0; int winAndStartedGame = 0; };
struct GameResults { int count = 0; int draw = 0; std::vector<PlayerGameResult> players = std::vector<PlayerGameResult>(2); };

int main()
{
    azb200::PlaySettings s;
    s.MCTS_SIMULATIONS = 16; s.THREADS_PER_MCTS = 2; s.BLOCKS = 2;
    if (az_device_count() == 0) {
        try { azb200::DevicePlay p(s); }
        catch (const std::runtime_error& e) { printf("NO_DEVICE_OK %s\n", e.what()); return 0; }
        printf("expected a failure without a GPU\n");
        return 1;
    }
    azb200::DevicePlay p(s);
    p.loadCheckpoint("/tmp/az_b200_play_ckpt.bin");
    GameResults a = p.playGames<GameResults>(41);          // odd request: 40 games, like Counter::hasNext(2)
    GameResults b = p.playGames<GameResults>(41);          // same seed, same weights: the match is reproducible
    bool ok = a.count == 40 && a.draw + a.players[0].win + a.players[1].win == 40 &&
              a.players[0].winAndStartedGame <= a.players[0].win && a.players[1].winAndStartedGame <= a.players[1].win &&
              b.count == a.count && b.draw == a.draw && b.players[0].win == a.players[0].win && b.players[1].win == a.players[1].win;
    printf("%s count=%d draw=%d az=%d/%d script=%d/%d\n", ok ? "PLAY_OK" : "PLAY_FAIL", a.count, a.draw, a.players[0].win,
           a.players[0].winAndStartedGame, a.players[1].win, a.players[1].winAndStartedGame);
    return ok ? 0 : 1;
}
