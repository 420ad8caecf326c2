// bridge/matcompile.cc -- command-line front end of the material compiler
// (refdrv_compile in matcompile_impl.icc):  ncb200_matcompile "<cfg-string>" out.ncb
// TEST INFRASTRUCTURE / reference-side tooling; links the unmodified reference.
#include <cstdio>
#include <cstdint>
#include <string>
extern "C" {
  void* refdrv_create( const char* cfg );
  void refdrv_destroy( void* );
  void* refdrv_compile_ex( void*, uint64_t* nbytes, unsigned flags );
  void refdrv_free( void* );
  const char* refdrv_lasterror();
}
int main( int argc, char** argv )
{
  // --vdos: S(alpha,beta) leaves derived from a phonon density of states are delivered as that density; the
  //         library expands them on the device (ncb_blob.h: NCB_KIND_SABVDOS)
  unsigned flags = 0;
  if ( argc == 4 && std::string(argv[1]) == "--vdos" ) { flags = 1; argv[1] = argv[2]; argv[2] = argv[3]; argc = 3; }
  if ( argc != 3 ) { std::fprintf(stderr,"usage: %s [--vdos] \"<cfg-string>\" out.ncb\n",argv[0]); return 2; }
  void* h = refdrv_create( argv[1] );
  if (!h) { std::fprintf(stderr,"error: %s\n",refdrv_lasterror()); return 1; }
  uint64_t n = 0;
  void* blob = refdrv_compile_ex( h, &n, flags );
  if (!blob) { std::fprintf(stderr,"error: %s\n",refdrv_lasterror()); return 1; }
  FILE* f = std::fopen( argv[2], "wb" );
  if (!f) { std::perror("fopen"); return 1; }
  std::fwrite( blob, 1, n, f );
  std::fclose(f);
  std::printf("%s: %llu bytes\n",argv[2],(unsigned long long)n);
  refdrv_free(blob);
  refdrv_destroy(h);
  return 0;
}
