/*
 * oracle/ref_kdtree_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A ctypes-friendly face over the REFERENCE'S OWN kd-tree (numcosmo/external/misc/kdtree.c + rb_knn_list.c, compiled
 * where they lie into oracle/_ref/libkdtree_ref.so by oracle/Makefile), used exactly as
 * _ncm_stats_dist_vkde_build_cov_array_kdtree does (ncm_stats_dist_vkde.c:374-389, 440-452): insert every point, rebuild,
 * k-nearest-neighbour search, walk the red-black list from its first element.  tests/test_oracle_kdtree.py checks the
 * oracle's brute-force (distance, index) ordering against it -- neighbour SET and ORDER, ties included.
 */
#include <stddef.h>
#include "kdtree.h"
#include "rb_knn_list.h"

/* points [n x d] row-major; for each of the nq query rows (indices into points) writes k neighbour indices and distances */
int
orc_ref_kdtree_knn (const double *points, int n, int d, const int *queries, int nq, int k, long *idx_out, double *dist_out)
{
  struct kdtree *tree = kdtree_init (d);
  int i, q;

  for (i = 0; i < n; i++)
    kdtree_insert (tree, (double *) &points[(size_t) i * d]);

  kdtree_rebuild (tree);

  for (q = 0; q < nq; q++)
  {
    rb_knn_list_table_t *table = kdtree_knn_search (tree, (double *) &points[(size_t) queries[q] * d], k);
    rb_knn_list_traverser_t trav;
    knn_list_t *p = rb_knn_list_t_first (&trav, table);
    int j = 0;

    do {
      if (j < k)
      {
        idx_out[(size_t) q * k + j]  = p->node->coord_index;
        dist_out[(size_t) q * k + j] = p->distance;
      }

      j++;
    } while ((p = rb_knn_list_t_next (&trav)) != NULL);

    rb_knn_list_destroy (table);

    if (j != k)
    {
      kdtree_destroy (tree);

      return -(q + 1);
    }
  }

  kdtree_destroy (tree);

  return 0;
}
