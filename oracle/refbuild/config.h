/* Hand-written stand-in for the autoconf-generated config.h of the reference
 * (configure.ac:4,18-19,71-76). Test infrastructure only: used to compile the
 * UNMODIFIED reference sources in /root/reference into oracle/_ref/. */
#ifndef ORACLE_REF_CONFIG_H
#define ORACLE_REF_CONFIG_H
#define HAVE_CLOCK_GETTIME 1
#define HAVE_OPENMP 1
#define HAVE_UNORDERED_MAP 1
#define HAVE_GOOGLE_SPARSE_HASH_MAP 1
#define PACKAGE_NAME "StriDe"
#define PACKAGE_VERSION "0.0.1"
#define PACKAGE_BUGREPORT "ythuang@cs.ccu.edu.tw"
#define PACKAGE_STRING "StriDe 0.0.1"
#endif
