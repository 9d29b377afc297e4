/* Stand-in for NumCosmo's meson-generated build_cfg.h, which numcosmo/external/levmar/levmar.h includes.
 * Empty on purpose: HAVE_LAPACK / HAVE_CONFIG_H stay undefined (see oracle/Makefile, target "ref"). */
#ifndef ORC_REF_BUILD_CFG_H
#define ORC_REF_BUILD_CFG_H
#endif
