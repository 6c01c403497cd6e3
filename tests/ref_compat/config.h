// config.h of the reference build (config.h.in): basic, fast tests
#ifndef CONFIG_H
#define CONFIG_H CONFIG_H
#define thorough_testing false
#endif
