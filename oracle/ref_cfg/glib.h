/* Stand-in for <glib.h> when compiling the reference's numcosmo/external/misc/kdtree.c and rb_knn_list.c in place
 * (oracle/Makefile, target "ref").  Those two files use nothing of GLib but the slice allocator and gint. */
#ifndef ORC_REF_GLIB_STUB_H
#define ORC_REF_GLIB_STUB_H
#include <stdlib.h>
typedef int gint;
#define g_slice_new(T) ((T *) malloc (sizeof (T)))
#define g_slice_new0(T) ((T *) calloc (1, sizeof (T)))
#define g_slice_free(T, p) free (p)
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#endif
