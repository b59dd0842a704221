#!/bin/bash
# usage: tools/try_lib.sh <lib.so> <cmd...> : run a command with an alternative build of the library
cp samurai_b200/libsamurai_b200.so /tmp/lib_backup.so
cp "$1" samurai_b200/libsamurai_b200.so
shift
"$@"
cp /tmp/lib_backup.so samurai_b200/libsamurai_b200.so
