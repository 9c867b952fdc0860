/* stand-in, see petscksp.h */
