// Test infrastructure only: stand-in header so the UNMODIFIED reference sources compile
// without Boost/autoconf (see oracle/README.md). Not part of the product.
#ifndef CONFIG_H
#define CONFIG_H 1
#define HAVE_STD_HASH 1
#define HAVE_UNORDERED_SET 1
#define HAVE_UNORDERED_MAP 1
#define HAVE_POPCNT 1
#define HAVE_GETOPT_LONG 1
#define HAVE_LIBDL 1
#define MAX_KMER 192
#define MAX_HASHES 32
#define FMBITS 64
#define PACKAGE_NAME "ABySS"
#define PACKAGE_BUGREPORT "abyss-users@bcgsc.ca"
#define VERSION "2.3.10"
#define PACKAGE_STRING "ABySS 2.3.10"
#endif
