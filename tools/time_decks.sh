#!/bin/bash
# wall time of the twin driver on the reference's FD decks (fd-64 ... fd-512: Re=1e4, alpha sweep, no shift guesses)
D=os-stab_b200/host/osb_drivers
T=$(mktemp -d)
cd $T
for spec in "64 0.02" "128 0.01" "256 0.01" "512 0.01"; do
  set -- $spec
  printf "$1\n10000\n0.76 1.16 $2\n0 0 1\neig-$1\n0\n" > deck.inp
  s=$(date +%s.%N)
  $OLDPWD/$D orrfdchan < deck.inp > out.txt 2>err.txt
  e=$(date +%s.%N)
  echo "N=$1 inc=$2 rows=$(grep -c E out.txt) wall=$(echo "$e - $s" | bc) s"
done
