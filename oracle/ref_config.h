/* Minimal stand-in for the autoconf-generated config.h of Jellyfish 2.3.0.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build).  We do not run the reference's
 * configure script; these are the handful of feature macros its sources test
 * (JF!include/jellyfish/mer_dna.hpp:26, rectangular_binary_matrix.hpp:142-159,
 * int128.hpp:14,119, generator_manager.hpp:26, locks_pthread.hpp:47). */
#ifndef ORACLE_REF_CONFIG_H
#define ORACLE_REF_CONFIG_H
#define HAVE_INT128 1
#define HAVE_NUMERIC_LIMITS128 1
#define HAVE_SI_INT 1
#if defined(__SSE2__)
#define HAVE_SSE 1
#define HAVE_SSE2 1
#endif
#define HAVE_EXT_STDIO_FILEBUF_H 1
#define HAVE_CLOCK_GETTIME 1
#define HAVE_EXECINFO_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_UNISTD_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_SYSCALL_H 1
#define HAVE_POSIX_MEMALIGN 1
#define PACKAGE "jellyfish"
#define PACKAGE_NAME "jellyfish"
#define PACKAGE_STRING "jellyfish 2.3.0"
#define PACKAGE_VERSION "2.3.0"
#define PACKAGE_BUGREPORT "gmarcais@umd.edu"
#define PACKAGE_URL ""
#define VERSION "2.3.0"
#endif
