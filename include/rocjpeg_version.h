/* Version macros of the rocJPEG C API this library is a drop-in for.
 * Mirrors the values in the reference's api/rocjpeg_version.h:35-37,49-52
 * (0.6.0) so that callers gated on ROCJPEG_CHECK_VERSION keep compiling. */
#ifndef ROCJPEG_VERSION_H
#define ROCJPEG_VERSION_H

#define ROCJPEG_MAJOR_VERSION 0
#define ROCJPEG_MINOR_VERSION 6
#define ROCJPEG_MICRO_VERSION 0

/* True when the library version is at least major.minor.micro. */
#define ROCJPEG_CHECK_VERSION(major, minor, micro)                                   \
    ((ROCJPEG_MAJOR_VERSION > (major)) ||                                            \
     (ROCJPEG_MAJOR_VERSION == (major) && ROCJPEG_MINOR_VERSION > (minor)) ||        \
     (ROCJPEG_MAJOR_VERSION == (major) && ROCJPEG_MINOR_VERSION == (minor) &&        \
      ROCJPEG_MICRO_VERSION >= (micro)))

#endif /* ROCJPEG_VERSION_H */
