"""Command-line shims with the reference's flag strings (histretch.cpp:68-74, aclahe.cpp:71-74, bgdehaze/main.py:24-30).
File I/O goes through cv2.imread / cv2.imwrite exactly like the reference; the pixels go through libuwip.so."""
