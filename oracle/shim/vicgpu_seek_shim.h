// oracle/shim/vicgpu_seek_shim.h -- force-included (after <iostream>) when read_vegparam.c / read_snowband.c are compiled for the
// drop-in's loader (oracle/Makefile, *_seek.o; vic_b200/host/vicgpu_fastread.h): the entry points get the names the indexed lookups
// call, and the files' one rewind() becomes the seek hook.  <cstdio> is included first because it #undefs rewind.
#include <cstdio>
#include <stdio.h>
extern "C" void vicgpu_reader_seek(FILE*);
#define rewind vicgpu_reader_seek
#define read_vegparam vicref_read_vegparam_at
#define read_snowband vicref_read_snowband_at
