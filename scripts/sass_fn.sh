#!/bin/bash
# usage: scripts/sass_fn.sh <lib.so> <substring of the mangled kernel name>  -> SASS of that function on stdout
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0} on {print}'
