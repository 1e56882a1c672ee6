extern "C" int oracle_shot_placeholder(void){return 0;}
