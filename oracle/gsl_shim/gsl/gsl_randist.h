/* TEST INFRASTRUCTURE ONLY (oracle/): the reference includes this header
 * (src/ratings.hh:14) but calls nothing from it. */
#ifndef HPF_SHIM_GSL_RANDIST_H
#define HPF_SHIM_GSL_RANDIST_H
#endif
