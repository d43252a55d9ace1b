g++ -O2 -std=c++17 -Iinclude tools/caller_bench.cpp -Lrovaca_b200 -lphmm_b200 -Wl,-rpath,'$ORIGIN/../rovaca_b200' -lpthread -o tools/caller_bench || exit 1
run() { echo "$1"; shift; env "$@" tools/caller_bench --threads 32 --regions 256 --depth 4 | cut -c1-260; }
run "default" X=1
run "poll 0" PHMM_B200_POLL_US=0
run "poll 50" PHMM_B200_POLL_US=50
run "poll 0 blocking sync" PHMM_B200_POLL_US=0 PHMM_B200_BLOCKING_SYNC=1
run "poll 50 blocking sync" PHMM_B200_POLL_US=50 PHMM_B200_BLOCKING_SYNC=1
echo "16 threads"; tools/caller_bench --threads 16 --regions 512 --depth 4 | cut -c1-260
echo "16 threads poll 0 blocking"; PHMM_B200_POLL_US=0 PHMM_B200_BLOCKING_SYNC=1 tools/caller_bench --threads 16 --regions 512 --depth 4 | cut -c1-260
echo "1 thread blocking"; PHMM_B200_BLOCKING_SYNC=1 tools/caller_bench --threads 1 --regions 512 | cut -c1-260
echo "128 threads default / blocking"; tools/caller_bench --threads 128 --regions 64 | cut -c1-200; PHMM_B200_POLL_US=0 PHMM_B200_BLOCKING_SYNC=1 tools/caller_bench --threads 128 --regions 64 | cut -c1-200
