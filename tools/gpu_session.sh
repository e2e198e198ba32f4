TREES=592 ROUNDS=2 tools/ab.sh base new noprefetch
