#!/bin/bash
# Type-check the C++ drop-in shims (and build the tiny host-logic test) against the stand-in OpenCV
# header; with a real OpenCV pass OPENCV_CFLAGS="$(pkg-config --cflags opencv4)" instead.
set -e
cd "$(dirname "$0")"
CFLAGS=${OPENCV_CFLAGS:--Iopencv_standin}
g++ -std=c++11 -Wall -Wextra -fsyntax-only $CFLAGS preprocessing_uwip.cpp
g++ -std=c++11 -Wall -Wextra -fsyntax-only -DUSE_GPU=1 $CFLAGS preprocessing_uwip.cpp
g++ -std=c++11 -Wall -O1 $CFLAGS shim_test.cpp preprocessing_uwip.cpp -L.. -luwip -Wl,-rpath,"$(cd .. && pwd)" -o shim_test
# the command-line shims (histretch.cpp:61-271, aclahe.cpp:64-226): built against the stand-in, whose imread / imwrite speak PPM
g++ -std=c++11 -Wall -O1 $CFLAGS histretch_main.cpp preprocessing_uwip.cpp -L.. -luwip -Wl,-rpath,"$(cd .. && pwd)" -o histretch
g++ -std=c++11 -Wall -O1 $CFLAGS aclahe_main.cpp -L.. -luwip -Wl,-rpath,"$(cd .. && pwd)" -o aclahe
echo "shims ok"
